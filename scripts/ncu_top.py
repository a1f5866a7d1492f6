"""Summarise an ncu report: key raw metrics + top stall lines.  usage: python scripts/ncu_top.py rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed", "launch__grid_size",
        "smsp__inst_executed.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic"]
for h, v in zip(hdr, vals):
    if h in want:
        print(f"{h:70s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print("total samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not" not in h]
agg = {}
for r in data:
    for h in stall_cols:
        agg[h] = agg.get(h, 0) + int(r[idx[h]] or 0)
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:n]:
    st = sorted(((h, int(r[idx[h]] or 0)) for h in stall_cols), key=lambda x: -x[1])[:2]
    print(r[idx["# Samples"]].rjust(6), r[idx.get("Address", 0)][-6:], r[idx["Source"]][:70].ljust(70), st)
