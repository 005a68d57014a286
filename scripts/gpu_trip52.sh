#!/bin/bash
# trip 52: fused pre-norm in the q/k/v GEMM: tests, decode microbench, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_golden.py -m gpu -q -x > gpurun_out/t52_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/t52_tests.log
timeout 300 python scripts/microbench_decode.py > gpurun_out/t52_micro.log 2>&1; tail -25 gpurun_out/t52_micro.log
timeout 600 python bench.py > gpurun_out/t52_bench.log 2>&1; tail -2 gpurun_out/t52_bench.log
