#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VF_TC_FIXED_BIAS=0.5 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_bwd_ops.py -q 2>&1 | tail -3
for b in 1.0 0.5; do
VF_TC_FIXED_BIAS=$b timeout 300 python scripts/layer_table.py > gpurun_out/layer_table_fixed_$b.txt 2>&1
echo "== bias $b"; grep "K\|{" gpurun_out/layer_table_fixed_$b.txt | awk '$2=="conv" && ($6==576 || ($5==192 && $6==192 && $7==1) || $6==960 || ($5==320&&$6==320&&$7==1) || ($5==64 && $7==1))' | head -12; tail -1 gpurun_out/layer_table_fixed_$b.txt
done
