#!/bin/bash
# round 2, run B: full GPU suite (margins), layer table, quick bench (no library baseline / extra configs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python scripts/layer_table.py > gpurun_out/layer_table_r02.txt 2>&1; tail -3 gpurun_out/layer_table_r02.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-library-baseline --no-extra-configs --no-full-generate --no-cpu-baseline > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; echo "bench rc=$?"
tail -c 2000 gpurun_out/bench_r02_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_b.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['kernel_classes'], d['train']['ms_per_step'], d['train']['value'])
PY
