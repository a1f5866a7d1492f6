#!/bin/bash
# GPU tests + the default bench line (argument: output tag, e.g. v6)
tag=${1:-tmp}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -1
timeout 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_r01_$tag.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r01_$tag.json"))
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["achieved"])
print(d["kernel_classes"])
print(d["train"]["ms_per_step"], d["train"]["value"], d["train"]["e2e"]["value"], d["cpu_baseline"]["value"], d["train"]["cpu_baseline"]["value"])
PY
