#!/bin/bash
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread"
python bench.py --ncu-minibatch > gpurun_out/plain_mb.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/mb_metrics_t72.csv python bench.py --ncu-minibatch > gpurun_out/ncu_mb.log 2>&1
echo "ncu mb exit $?"
