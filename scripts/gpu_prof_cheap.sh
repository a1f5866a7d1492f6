#!/bin/bash
# A profile set that costs ~3 GPU-minutes instead of 30 (scripts/gpu_prof_r01_v6.sh replays 227 launches ~40x each):
#   * launch lists (gpu__time_duration only, one replay) of a sampling run and a training step,
#   * ONE --set full capture each of five representative convolutions (prof_conv.py cases) with SASS-level stall samples,
#   * --set full of 12 consecutive non-conv launches of a sampling step (GroupNorm, attention, up-sampling, compose).
# usage: bash scripts/gpu_prof_cheap.sh <tag>
tag=${1:-rNN}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches_sampling.csv $B --no-train > gpurun_out/ncu_l1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1600 --csv --log-file gpurun_out/${tag}_launches_train.csv python scripts/train_probe.py 28 noprof > gpurun_out/ncu_l2.log 2>&1
for c in c64 u64 c128 c192 c320; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 1 -o gpurun_out/${tag}_conv_$c -f python scripts/prof_conv.py $c typical > gpurun_out/ncu_$c.log 2>&1
  python scripts/ncu_top.py gpurun_out/${tag}_conv_$c.ncu-rep 25 > gpurun_out/${tag}_ncu_conv_${c}_source.txt 2>&1
done
T=/tmp/vfprof; mkdir -p $T
timeout 300 ncu --set full --clock-control none -k regex:"gn_apply_kernel|compose|attn_tc|upsample" -s 100 -c 12 -o $T/bw -f $B --no-train > gpurun_out/ncu_bw.log 2>&1
python scripts/ncu_summary.py $T/bw.ncu-rep > gpurun_out/${tag}_ncu_full_bw.txt 2>&1
ls -la gpurun_out | tail -15
