#!/bin/bash
# round 2 evidence: launch lists (sampling, training) and a counter capture of the 84 convolution launches of one sampling step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 172 -c 344 --csv --log-file gpurun_out/r02_launches_sampling.csv python scripts/sample_launches.py 3 > gpurun_out/ncu_s.log 2>&1; tail -1 gpurun_out/ncu_s.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_train.csv python scripts/train_launches.py 3 > gpurun_out/ncu_t.log 2>&1; tail -1 gpurun_out/ncu_t.log
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_bytes.sum --clock-control none -k regex:conv_tc -s 84 -c 84 -o gpurun_out/prof_conv_r02 -f python scripts/sample_launches.py 2 > gpurun_out/ncu_c.log 2>&1; tail -2 gpurun_out/ncu_c.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_*.csv
