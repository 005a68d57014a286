#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/t70_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/t70_tests.log
timeout 300 python scripts/microbench_glu_train.py > gpurun_out/t70_micro.log 2>&1; tail -6 gpurun_out/t70_micro.log | head -4
timeout 300 python scripts/microbench_decode.py > gpurun_out/t70_microd.log 2>&1; grep "glu_fused\|decode_step" gpurun_out/t70_microd.log | head -3
timeout 600 python bench.py > gpurun_out/t70_bench.log 2>&1; tail -1 gpurun_out/t70_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['phases_ms'], d['roofline']['frac'])"
