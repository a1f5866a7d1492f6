#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,lts__t_sectors_op_red.sum --clock-control none -k regex:wgrad -c 200 --csv --log-file gpurun_out/wgrad_launches.csv python scripts/train_launches.py 2 > gpurun_out/ncu_wg.log 2>&1
tail -2 gpurun_out/ncu_wg.log
