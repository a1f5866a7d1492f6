#!/bin/bash
# round 2, run D: streaming GroupNorm kernels: tests, A/B bench (VF_GN_STREAM=0/1), training launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for v in 0 1; do
VF_GN_STREAM=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-library-baseline --no-extra-configs --no-full-generate --no-cpu-baseline > gpurun_out/bench_r02_d$v.json 2> gpurun_out/bench_r02_d$v.err; echo "bench GN_STREAM=$v rc=$?"; tail -c 600 gpurun_out/bench_r02_d$v.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02_d$v.json'))
print("GN_STREAM=$v", {k:d[k] for k in ('value','ms_per_step')}, d['kernel_classes']['gn_apply'], 'conv', d['kernel_classes']['conv']['ms_per_step'], 'train ms', d['train']['ms_per_step'])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train_r02.csv python scripts/train_launches.py 2 > gpurun_out/ncu_train.log 2>&1
python scripts/ncu_classes.py gpurun_out/launches_train_r02.csv | head -30
