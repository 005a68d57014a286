#!/bin/bash
# trip 14: fused obs encoder bring-up (descriptor order), fused policy head; full tests + bench + decode-step metrics
mkdir -p gpurun_out
rm -f gpurun_out/t14.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "obs_embed_fused" > gpurun_out/t14_obs0.log 2>&1
echo "obs swap=0 exit $?" >> gpurun_out/t14.log
MPX_OBS_SWAP=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "obs_embed_fused" > gpurun_out/t14_obs1.log 2>&1
echo "obs swap=1 exit $?" >> gpurun_out/t14.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t14_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t14.log
tail -8 gpurun_out/t14_tests.log >> gpurun_out/t14.log
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/t14_bench.log 2>&1
echo "bench exit $?" >> gpurun_out/t14.log
tail -2 gpurun_out/t14_bench.log >> gpurun_out/t14.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread"
python bench.py --ncu-decode-step > gpurun_out/plain_ds.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/ds_metrics_t14.csv python bench.py --ncu-decode-step > gpurun_out/ncu_ds.log 2>&1
echo "ncu ds exit $?" >> gpurun_out/t14.log
cat gpurun_out/t14.log
