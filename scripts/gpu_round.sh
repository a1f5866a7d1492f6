#!/bin/bash
# Runs the GPU test groups in separate processes (a trapped kernel poisons only its own group).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 python -m pytest "$@" -q --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/$name.log; }
run ops_safe tests/test_gpu_ops.py -k "not tcgen05 and not bf16 and not attention"
run ops_gn_bf16 tests/test_gpu_ops.py -k "groupnorm"
run conv_tc tests/test_gpu_ops.py -k "test_conv_bf16_tcgen05"
run conv_tc_misc tests/test_gpu_ops.py -k "final_layer or qkv_split"
run attn tests/test_gpu_ops.py -k "test_attention"
run unet_fp32 tests/test_gpu_unet.py -k "fp32 or generate"
run unet_bf16 tests/test_gpu_unet.py -k "bf16"
