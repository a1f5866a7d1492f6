timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -1
timeout 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_r01_v5.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r01_v5.json"))
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["achieved"])
print(d["kernel_classes"])
print(d["train"]["ms_per_step"], d["train"]["value"], d["train"]["e2e"]["value"], d["cpu_baseline"]["value"], d["train"]["cpu_baseline"]["value"])
PY
