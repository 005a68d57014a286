#!/bin/bash
# trip 93: end-of-round decode-step kernel list (DRAM traffic) and one --set full capture of the roofline kernel
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread"
python bench.py --ncu-decode-step > gpurun_out/t93_plain_ds.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/ds_metrics_t93.csv python bench.py --ncu-decode-step > gpurun_out/t93_ncu_ds.log 2>&1
echo "ncu ds exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_decode_fused_kernel -c 1 -o gpurun_out/attn_decode_fused_t93 -f python bench.py --ncu-decode-step > gpurun_out/t93_ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/attn_decode_fused_t93.ncu-rep gpurun_out/ds_metrics_t93.csv
