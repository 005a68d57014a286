#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "glu_bwd_fused" > gpurun_out/t54_tests.log 2>&1; echo "tests exit $?"; tail -15 gpurun_out/t54_tests.log
timeout 300 python scripts/stamps_glu_train.py > gpurun_out/t54_stamps.log 2>&1; cat gpurun_out/t54_stamps.log
