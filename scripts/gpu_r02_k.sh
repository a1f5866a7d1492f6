#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_fullsize.py -q 2>&1 | tail -5
for v in 0 1; do
VF_WG64=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-library-baseline --no-extra-configs --no-full-generate --no-cpu-baseline --no-strong --train-steps 20 2>/dev/null | tail -1 > gpurun_out/bench_wg64_$v.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_wg64_$v.json').read().strip().splitlines()[-1])
print("VF_WG64=$v train ms", d['train']['ms_per_step'], d['train']['value'])
PY
done
