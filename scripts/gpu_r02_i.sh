#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for acc in 1 2 3 4; do
VF_WG_ACC_SMALL=$acc timeout 600 python bench.py --steps 5 --warmup 3 --no-library-baseline --no-extra-configs --no-full-generate --no-cpu-baseline --no-strong --train-steps 20 2>/dev/null | tail -1 > gpurun_out/bench_wg$acc.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_wg$acc.json').read().strip().splitlines()[-1])
print("VF_WG_ACC_SMALL=$acc train ms", d['train']['ms_per_step'])
PY
done
