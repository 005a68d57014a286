#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t63_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/t63_tests.log
timeout 600 python bench.py > gpurun_out/t63_bench.log 2>&1; tail -1 gpurun_out/t63_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['phases_ms'], d['roofline']['frac'], d['roofline']['avg_launch_us'])"
timeout 300 python scripts/microbench_decode.py > gpurun_out/t63_micro.log 2>&1; tail -16 gpurun_out/t63_micro.log | head -14
