#!/bin/bash
# First GPU trip: each kernel family in its own process so one fault does not hide the rest.
mkdir -p gpurun_out
rm -f gpurun_out/probe.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for k in gae hlgauss linear_plain linear_epilogues linear_wgrad qkv_rope glu_up attn_decode; do
  echo "=== $k" >> gpurun_out/t1.log
  timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "$k" >> gpurun_out/t1.log 2>&1
  echo "exit $?" >> gpurun_out/t1.log
done
tail -5 gpurun_out/t1.log
