#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
GN_SHAPES=1 GN_ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_stream -s 6 -c 3 -o gpurun_out/prof_gn_stream -f python scripts/gn_bench.py > gpurun_out/ncu_gn_stream.log 2>&1
tail -3 gpurun_out/ncu_gn_stream.log
GN_SHAPES=1 GN_ITERS=1 VF_GN_STREAM=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_ -s 6 -c 3 -o gpurun_out/prof_gn_old -f python scripts/gn_bench.py > gpurun_out/ncu_gn_old.log 2>&1
tail -3 gpurun_out/ncu_gn_old.log
ls -la gpurun_out/*.ncu-rep
