#!/bin/bash
# trip 53: fused GLU backward: parity test + training-size microbench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "glu_bwd_fused" > gpurun_out/t53_tests.log 2>&1; echo "tests exit $?"; tail -15 gpurun_out/t53_tests.log
timeout 300 python scripts/microbench_glu_train.py > gpurun_out/t53_micro.log 2>&1; tail -6 gpurun_out/t53_micro.log
