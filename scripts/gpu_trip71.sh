#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t71_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/t71_tests.log
timeout 300 python scripts/microbench_decode.py > gpurun_out/t71_microd.log 2>&1; grep "us " gpurun_out/t71_microd.log | head -14
timeout 600 python bench.py > gpurun_out/t71_bench.log 2>&1; tail -1 gpurun_out/t71_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['phases_ms'], d['roofline']['frac'])"
