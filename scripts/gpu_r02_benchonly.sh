#!/bin/bash
# the two bench arms of the final state -> gpurun_out/r02_bench.json, r02_bench_reference_arm.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final.out 2> gpurun_out/bench_r02_final.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_r02_final.out > gpurun_out/r02_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference_arm.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "gen_full", d["generate_full"]["value"],
      "train", d["train"]["ms_per_step"], d["train"]["value"], d["train"]["e2e"]["value"], "strong", d["train"]["strong"]["value"])
r = json.load(open("gpurun_out/r02_bench_reference_arm.json"))
print("reference arm", r["value"], r.get("cpu_baseline", {}).get("cores"))
PY
