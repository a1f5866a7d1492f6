#!/bin/bash
# round 2, run A: full GPU test suite (with recorded margins), smoke, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_r02_a.err
head -c 6000 gpurun_out/bench_r02_a.json
