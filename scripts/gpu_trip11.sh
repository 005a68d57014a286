#!/bin/bash
# trip 11: coalesced GEMM epilogue, parallel wgrad reduce, golden fixtures; per-kernel metrics + full ncu of GEMM kernels
mkdir -p gpurun_out
rm -f gpurun_out/t11.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t11_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t11.log
tail -5 gpurun_out/t11_tests.log >> gpurun_out/t11.log
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/t11_bench.log 2>&1
echo "bench exit $?" >> gpurun_out/t11.log
tail -2 gpurun_out/t11_bench.log >> gpurun_out/t11.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread"
python bench.py --ncu-minibatch > gpurun_out/plain_mb.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/mb_metrics_t11.csv python bench.py --ncu-minibatch > gpurun_out/ncu_mb.log 2>&1
echo "ncu mb exit $?" >> gpurun_out/t11.log
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base demangled \
  -k 'regex:.*gemm_tn_kernel<(256|128), (0|2)>.*' -c 8 -f -o gpurun_out/gemm_full_t11 python bench.py --ncu-minibatch > gpurun_out/ncu_gemm_full.log 2>&1
echo "ncu gemm full exit $?" >> gpurun_out/t11.log
ls -la gpurun_out/*.ncu-rep >> gpurun_out/t11.log 2>&1
cat gpurun_out/t11.log
