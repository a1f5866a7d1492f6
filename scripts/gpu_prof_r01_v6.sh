#!/bin/bash
# Round-1 (v6) profile set: launch lists of one sampling run and one training step, plus --set full captures of the hot
# kernels.  Reports are summarised on the GPU box (gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_launches_v6_sampling.csv $B --no-train > gpurun_out/ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1600 --csv --log-file gpurun_out/r01_launches_v6_train.csv python scripts/train_probe.py 28 noprof > gpurun_out/ncu_l2.log 2>&1
T=/tmp/vfprof; mkdir -p $T
timeout 900 ncu --set full --clock-control none -k regex:"conv_tc_kernel" -s 84 -c 84 -o $T/conv -f $B --no-train > gpurun_out/ncu_f1.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"gn_apply_kernel|compose|pack_views|attn_tc|embed|upsample" -s 83 -c 83 -o $T/bw -f $B --no-train > gpurun_out/ncu_f2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"wgrad_tc|attn_bwd|gn_bwd|colsum|compose_mse" -s 300 -c 60 -o $T/train -f python scripts/train_probe.py 28 noprof > gpurun_out/ncu_f3.log 2>&1
for n in conv bw train; do python scripts/ncu_summary.py $T/$n.ncu-rep > gpurun_out/r01_ncu_full_v6_$n.txt 2>&1; done
ls -la gpurun_out/ $T | tail -20
