#!/bin/bash
# one ncu --set full capture (with SASS-level stall samples) of the 64x64 64->64 convolution, typical (no residual) variant
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 1 -o gpurun_out/prof_c64t -f python scripts/prof_conv.py c64 typical > gpurun_out/ncu_c64t.log 2>&1
tail -2 gpurun_out/ncu_c64t.log
for c in c64 u64 c128; do python scripts/prof_conv.py $c typical; done > gpurun_out/prof_v4red.txt 2>&1
