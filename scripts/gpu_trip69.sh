#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "glu" > gpurun_out/t69_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/t69_tests.log
timeout 300 python scripts/microbench_glu_train.py > gpurun_out/t69_micro.log 2>&1; tail -6 gpurun_out/t69_micro.log | head -4
timeout 300 python scripts/stamps_glu_train.py > gpurun_out/t69_stamps.log 2>&1; grep -A40 "glu_bwd_fused" gpurun_out/t69_stamps.log | grep "slot" | head -32
