#!/bin/bash
# round 2: ncu --set full captures of the hot kernels (one sampling step's convolutions, its bandwidth kernels, the backward kernels of
# one training step), summarised on the GPU box (the reports are too large to bring back)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=/tmp/vfprof; mkdir -p $T
timeout 1200 ncu --set full --clock-control none -k regex:"conv_tc_kernel" -s 84 -c 84 -o $T/conv -f python scripts/sample_launches.py 2 > gpurun_out/ncu_f1.log 2>&1; tail -1 gpurun_out/ncu_f1.log
timeout 900 ncu --set full --clock-control none -k regex:"gn_apply_kernel|compose|pack_views|attn_tc|embed|upsample|step_prepare" -s 85 -c 85 -o $T/bw -f python scripts/sample_launches.py 2 > gpurun_out/ncu_f2.log 2>&1; tail -1 gpurun_out/ncu_f2.log
timeout 1200 ncu --set full --clock-control none -k regex:"wgrad|attn_bwd|gn_bwd|colsum|compose_mse|adam" -s 190 -c 190 -o $T/train -f python scripts/train_launches.py 2 > gpurun_out/ncu_f3.log 2>&1; tail -1 gpurun_out/ncu_f3.log
for n in conv bw train; do python scripts/ncu_summary.py $T/$n.ncu-rep > gpurun_out/r02_ncu_full_$n.txt 2>&1; done
ls -la $T | tail -5; wc -l gpurun_out/r02_ncu_full_*.txt
