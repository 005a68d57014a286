#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "glu" > gpurun_out/t55_tests.log 2>&1; echo "tests exit $?"; tail -15 gpurun_out/t55_tests.log
timeout 300 python scripts/microbench_glu_train.py > gpurun_out/t55_micro.log 2>&1; tail -6 gpurun_out/t55_micro.log
timeout 300 python scripts/stamps_glu_train.py > gpurun_out/t55_stamps.log 2>&1; cat gpurun_out/t55_stamps.log
