#!/bin/bash
# trip 89: end-of-round evidence: GPU suite, smoke, default bench line (with cpu_baseline), reference arm, minibatch kernel table
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t89_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/t89_tests.log
timeout 300 python scripts/run_smoke.py > gpurun_out/t89_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/t89_smoke.log
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/t89_bench.log 2> gpurun_out/t89_bench.err; echo "bench exit $?"; tail -1 gpurun_out/t89_bench.log | cut -c1-2500
grep -E "Elapsed|Maximum resident" gpurun_out/t89_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/t89_ref.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/t89_ref.log | cut -c1-900
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread"
python bench.py --ncu-minibatch > gpurun_out/t89_plain_mb.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/mb_metrics_t89.csv python bench.py --ncu-minibatch > gpurun_out/t89_ncu_mb.log 2>&1
echo "ncu mb exit $?"
