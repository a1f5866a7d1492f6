#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json | cut -c1-400
tail -3 gpurun_out/bench_err.log
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/bench_ref.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 150 -c 2 -o gpurun_out/prof_conv_bench -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 150 -c 1 -o gpurun_out/prof_gn_bench -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_gn.log 2>&1
ls -la gpurun_out/*.ncu-rep
