#!/bin/bash
# trip 23: full GPU suite, default bench (with cpu_baseline), reference arm, smoke
mkdir -p gpurun_out
rm -f gpurun_out/t23.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t23_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t23.log
tail -4 gpurun_out/t23_tests.log >> gpurun_out/t23.log
timeout 900 python bench.py > gpurun_out/t23_bench_default.log 2>&1
echo "bench default exit $?" >> gpurun_out/t23.log
tail -1 gpurun_out/t23_bench_default.log >> gpurun_out/t23.log
timeout 900 python bench.py --impl reference > gpurun_out/t23_bench_reference.log 2>&1
echo "bench reference exit $?" >> gpurun_out/t23.log
tail -1 gpurun_out/t23_bench_reference.log >> gpurun_out/t23.log
python scripts/run_smoke.py > gpurun_out/t23_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/t23.log
tail -2 gpurun_out/t23_smoke.log >> gpurun_out/t23.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread"
python bench.py --ncu-minibatch > gpurun_out/plain_mb.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/mb_metrics_t23.csv python bench.py --ncu-minibatch > gpurun_out/ncu_mb.log 2>&1
echo "ncu mb exit $?" >> gpurun_out/t23.log
cat gpurun_out/t23.log
