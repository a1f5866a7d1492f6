#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/prof_c0 -f python scripts/prof_conv.py c0 > gpurun_out/ncu_c0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/prof_c128 -f python scripts/prof_conv.py c128 > gpurun_out/ncu_c128.log 2>&1
tail -1 gpurun_out/ncu_c128.log
