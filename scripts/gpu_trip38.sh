#!/bin/bash
# trip 38: launch list of ONE bench step (one PPO update) under ncu, + full GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t38_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t38_tests.log
python bench.py --ncu > gpurun_out/plain_step.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_launches_t38.csv python bench.py --ncu > gpurun_out/ncu_step.log 2>&1
echo "ncu step exit $?"; ls -la gpurun_out/step_launches_t38.csv
