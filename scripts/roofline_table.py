"""Per-launch roofline table from a scripts/layer_table.py dump (no GPU needed):
python scripts/roofline_table.py profiles/r01_layer_table_v8.txt > profiles/r01_roofline_per_layer_v8.txt

For every convolution and GroupNorm launch: measured us (CUDA events, PDL suspended, so each figure carries ~3-4 us of
launch + dependency latency), the time the tensor pipe would need at the measured sustained bf16 peak, the time HBM would
need for the algorithmic bytes at the measured copy bandwidth, and measured / max(of the two) = distance to the roofline."""
import json
import sys

PEAK = json.load(open("MEASURED_PEAKS.json"))
TF, GBS = PEAK["bf16_tflops_sustained"], PEAK["hbm_gbs"]

rows = []
for line in open(sys.argv[1]):
    f = line.split()
    if len(f) < 9 or not f[0].isdigit():
        continue
    idx, kind, img, H, cin, cout, ks, st, us = int(f[0]), f[1], int(f[2]), int(f[3]), int(f[4]), int(f[5]), int(f[6]), int(f[7]), float(f[8])
    if kind == "conv" and H:
        Ho = H // 2 if st == 2 else H
        flops = 2.0 * img * Ho * Ho * cin * cout                       # cin column holds the total K of the launch
        kk = ks * ks
        c_in = cin / kk if kk else cin                                  # approximate input channels (extra 1x1 segments counted as 3x3 taps: upper bound on K only)
        byts = 2.0 * img * (H * H * c_in + Ho * Ho * cout) + 2.0 * cin * cout
        t_tc, t_hbm = flops / (TF * 1e6), byts / (GBS * 1e3)
        rows.append((idx, kind, f"{H}x{H} K={cin} N={cout} k{ks} s{st}", us, t_tc, t_hbm))
    elif kind == "gn_apply" and H:
        byts = 4.0 * img * H * H * cin
        rows.append((idx, kind, f"{H}x{H} C={cin}", us, 0.0, byts / (GBS * 1e3)))

print(f"# peaks: {TF} TFLOP/s sustained bf16, {GBS} GB/s HBM copy (MEASURED_PEAKS.json); times in us; source {sys.argv[1]}")
print(f"{'#':>3s} {'class':9s} {'shape':28s} {'meas':>7s} {'tensor':>7s} {'hbm':>7s} {'bound':>6s} {'meas/roof':>9s}")
tot = {}
for idx, kind, shape, us, t_tc, t_hbm in rows:
    roof = max(t_tc, t_hbm)
    print(f"{idx:3d} {kind:9s} {shape:28s} {us:7.1f} {t_tc:7.1f} {t_hbm:7.1f} {'tensor' if t_tc >= t_hbm else 'hbm':>6s} {us / roof:9.2f}")
    a = tot.setdefault(kind, [0.0, 0.0])
    a[0] += us; a[1] += roof
for k, (m, r) in tot.items():
    print(f"# {k}: measured {m / 1e3:.3f} ms per step, roofline {r / 1e3:.3f} ms -> {r / m:.1%} of the roofline")
