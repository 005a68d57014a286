#!/bin/bash
# trip 80: shared-memory row assembler for the observation layout change; whole-step ncu launch list of the current code
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t80_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t80_tests.log
timeout 300 python bench.py --skip-cpu --steps 10 --warmup 3 > gpurun_out/t80_bench.log 2>&1
echo -n "bench: "; tail -1 gpurun_out/t80_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), round(d['e2e']['value']/1e6,3), d['phases_ms'], d['roofline']['frac'])"
python bench.py --ncu > gpurun_out/t80_plain_step.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_launches_t80.csv python bench.py --ncu > gpurun_out/t80_ncu_step.log 2>&1
echo "ncu step exit $?"; ls -la gpurun_out/step_launches_t80.csv
