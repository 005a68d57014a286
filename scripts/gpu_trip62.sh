#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t62_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/t62_tests.log
timeout 300 python scripts/microbench_attn_decode.py --fused > gpurun_out/t62_attn_fused.log 2>&1; cat gpurun_out/t62_attn_fused.log
timeout 600 python bench.py > gpurun_out/t62_bench.log 2>&1; tail -1 gpurun_out/t62_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['phases_ms'], d['roofline']['frac'], d['roofline']['avg_launch_us'])"
