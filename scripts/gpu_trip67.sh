#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t67_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/t67_tests.log
timeout 300 python scripts/microbench_glu_train.py > gpurun_out/t67_micro.log 2>&1; tail -6 gpurun_out/t67_micro.log | head -4
timeout 300 python scripts/stamps_glu_train.py > gpurun_out/t67_stamps.log 2>&1; grep -A60 "glu_fused fwd" gpurun_out/t67_stamps.log | grep "slot" | head -50 | awk 'NR>=4 && NR<=22 || NR>=36 && NR<=47'
timeout 600 python bench.py > gpurun_out/t67_bench.log 2>&1; tail -1 gpurun_out/t67_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['phases_ms'], d['roofline']['frac'])"
