#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t74_tests.log 2>&1; echo "tests exit $?"; tail -6 gpurun_out/t74_tests.log
timeout 600 python bench.py --skip-cpu > gpurun_out/t74_bench.log 2>&1; tail -1 gpurun_out/t74_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['phases_ms'])"
timeout 300 python scripts/microbench_gemm.py > gpurun_out/t74_gemm.log 2>&1; tail -25 gpurun_out/t74_gemm.log
