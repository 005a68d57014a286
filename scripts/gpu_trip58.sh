#!/bin/bash
# trip 58: profiles after the fused decode attention + fused GLU training kernels:
#  decode-step kernel list with DRAM traffic, minibatch kernel list, one --set full capture of the decode attention
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread"
python bench.py --ncu-decode-step > gpurun_out/plain_ds.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/ds_metrics_t58.csv python bench.py --ncu-decode-step > gpurun_out/ncu_ds.log 2>&1
echo "ncu ds exit $?"
python bench.py --ncu-minibatch > gpurun_out/plain_mb.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/mb_metrics_t58.csv python bench.py --ncu-minibatch > gpurun_out/ncu_mb.log 2>&1
echo "ncu mb exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_decode_fused_kernel -c 1 -o gpurun_out/attn_decode_fused_t58 -f python bench.py --ncu-decode-step > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep gpurun_out/*_t58.csv
timeout 600 python bench.py > gpurun_out/t58_bench.log 2>&1; tail -1 gpurun_out/t58_bench.log
