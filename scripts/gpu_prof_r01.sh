#!/bin/bash
# Round-1 profile set: launch lists of one sampling run and one training step, plus --set full captures of the hot kernels.
# The reports are summarised on the GPU box (gpurun_out/ is capped at 64 MiB); only one small report is kept.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_launches_v4_sampling.csv $B --no-train > gpurun_out/ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1600 --csv --log-file gpurun_out/r01_launches_v4_train.csv python scripts/train_probe.py 28 noprof > gpurun_out/ncu_l2.log 2>&1
T=/tmp/vfprof; mkdir -p $T
timeout 900 ncu --set full --clock-control none -k regex:"conv_tc_kernel" -s 84 -c 84 -o $T/conv -f $B --no-train > gpurun_out/ncu_f1.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"gn_apply_kernel|compose|pack_views|attn_tc|embed|upsample" -s 86 -c 86 -o $T/bw -f $B --no-train > gpurun_out/ncu_f2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"wgrad_tc|attn_bwd|gn_bwd|colsum|compose_mse" -s 330 -c 60 -o $T/train -f python scripts/train_probe.py 28 noprof > gpurun_out/ncu_f3.log 2>&1
for n in conv bw train; do python scripts/ncu_summary.py $T/$n.ncu-rep > gpurun_out/r01_ncu_full_$n.txt 2>&1; done
# one source-level report of a representative convolution (32x32, 128 -> 128) for the record
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/r01_conv_c128 -f python scripts/prof_conv.py c128 > gpurun_out/ncu_f4.log 2>&1
ls -la gpurun_out/ $T
