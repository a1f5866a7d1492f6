"""Per-launch key metrics of an ncu report (raw page): python scripts/ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv, io, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kname = col.get("Kernel Name")
    print(f"# {rep}: {len(rows) - 2} launches (ncu --set full --clock-control none; per-launch, cold-cache, serialised)")
    for r in rows[2:]:
        name = r[kname].split("(")[0].replace("void ", "")
        parts = []
        for key, short in WANT:
            if key in col:
                parts.append(f"{short}={r[col[key]]}{units[col[key]] and ' ' + units[col[key]]}")
        print(f"{name}\n    " + "  ".join(parts))


if __name__ == "__main__":
    main()
