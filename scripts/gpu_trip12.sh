#!/bin/bash
# trip 12: golden GPU tests + full ncu (source-level) of the epilogue-heavy GEMMs
mkdir -p gpurun_out
rm -f gpurun_out/t12.log
timeout 600 python -m pytest tests/test_gpu_golden.py -m gpu -q > gpurun_out/t12_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t12.log
tail -3 gpurun_out/t12_tests.log >> gpurun_out/t12.log
python bench.py --ncu-minibatch > gpurun_out/plain_mb.log 2>&1 && \
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base demangled \
  -k 'regex:.*gemm_tn_kernel<\(int\)256, \(int\)(0|2)>.*' -c 3 -f -o gpurun_out/gemm_full_t12 python bench.py --ncu-minibatch > gpurun_out/ncu_gemm_full.log 2>&1
echo "ncu gemm full exit $?" >> gpurun_out/t12.log
ls -la gpurun_out/*.ncu-rep >> gpurun_out/t12.log 2>&1
cat gpurun_out/t12.log
