#!/bin/bash
# Runs the GPU test groups in separate processes (a trapped kernel poisons only its own group).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 python -m pytest "$@" -q --tb=short -p no:cacheprovider -s > gpurun_out/$name.log 2>&1; echo "exit $?"; grep -E "UMMA_SHIFT_PROBE|passed|failed|Error|error" gpurun_out/$name.log | tail -n 12; }
run probe tests/test_gpu_ops.py -k "probe"
run ops tests/test_gpu_ops.py -k "not probe"
run unet tests/test_gpu_unet.py
if [ "$1" == "bench" ]; then bash scripts/gpu_bench.sh; fi
