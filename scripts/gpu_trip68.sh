#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "glu" > gpurun_out/t68_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/t68_tests.log
timeout 300 python scripts/microbench_glu_train.py > gpurun_out/t68_micro.log 2>&1; tail -6 gpurun_out/t68_micro.log | head -4
timeout 300 python scripts/stamps_glu_train.py > gpurun_out/t68_stamps.log 2>&1; grep -A60 "glu_fused fwd" gpurun_out/t68_stamps.log | grep "slot" | head -50 | awk 'NR>=4 && NR<=22 || NR>=36 && NR<=47'
timeout 300 python scripts/microbench_decode.py > gpurun_out/t68_microd.log 2>&1; grep "glu_fused\|decode_step" gpurun_out/t68_microd.log | head -4
