"""Per-kernel-name totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python scripts/ncu_classes.py file.csv [launches_to_skip]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot, cnt = collections.Counter(), collections.Counter()
for i, r in enumerate(rd):
    if i < skip:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    n = re.sub(r"\(.*", "", r[ki])
    n = re.sub(r"<.*", "", n).split("::")[-1] or r[ki][:40]
    tot[n] += v
    cnt[n] += 1
s = sum(tot.values())
print(f"# {sys.argv[1]}: {sum(cnt.values())} launches, {s / 1e6:.3f} ms (cold-cache, serialised)")
for n, v in tot.most_common(40):
    print(f"{n:44s} {v / 1e3:10.1f} us {cnt[n]:5d} launches {100 * v / s:5.1f} %")
