"""Summarise an ncu capture of the convolution launches of ONE sampling step:
python scripts/ncu_conv_summary.py gpurun_out/prof_conv_r02.ncu-rep profiles/r02_ncu_conv.txt profiles/conv_traffic.json"""
import csv, io, json, subprocess, sys
rep, out_txt, out_json = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = csv.reader(io.StringIO(raw))
hdr = next(rd); next(rd)
col = {n: i for i, n in enumerate(hdr)}
def g(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return default
rows = [r for r in rd if "conv_tc_kernel" in r[col["Kernel Name"]]]
t = [g(r, "gpu__time_duration.sum") for r in rows]                   # us
rd_b = [g(r, "dram__bytes_read.sum") for r in rows]                  # MB (ncu scales units per column)
wr_b = [g(r, "dram__bytes_write.sum") for r in rows]
tp = [g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") for r in rows]
unit_r = "MB"
tot_t = sum(t)
w_tp = sum(a * b for a, b in zip(t, tp)) / tot_t if tot_t else 0.0
lines = [f"# {rep}: {len(rows)} conv_tc_kernel launches, {tot_t / 1e3:.3f} ms summed (cold-cache, serialised)",
         f"# tensor pipe active (time-weighted): {w_tp:.1f} %   DRAM read {sum(rd_b) / 1e3:.3f} GB  written {sum(wr_b) / 1e3:.3f} GB",
         f"{'#':>3s} {'us':>8s} {'tensor%':>8s} {'rd MB':>8s} {'wr MB':>8s}  grid"]
for i, r in enumerate(rows):
    lines.append(f"{i:3d} {t[i]:8.1f} {tp[i]:8.1f} {rd_b[i]:8.1f} {wr_b[i]:8.1f}  {r[col['Grid Size']] if 'Grid Size' in col else ''}")
open(out_txt, "w").write("\n".join(lines) + "\n")
json.dump({"bytes_per_launch": (sum(rd_b) + sum(wr_b)) * 1e6 / max(1, len(rows)), "launches": len(rows),
           "dram_read_bytes_per_step": sum(rd_b) * 1e6, "dram_write_bytes_per_step": sum(wr_b) * 1e6,
           "tensor_pipe_pct_time_weighted": w_tp,
           "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over the {len(rows)} conv_tc_kernel launches of one sampling step "
                     f"({out_txt}), per launch"}, open(out_json, "w"), indent=1)
print("\n".join(lines[:3]))
