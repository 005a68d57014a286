#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/t6.log gpurun_out/probe.log
run() { echo "=== $*" >> gpurun_out/t6.log; timeout 600 "$@" >> gpurun_out/t6.log 2>&1; echo "exit $?" >> gpurun_out/t6.log; }
for k in "attn_decode" "glu_fused" "not attn_decode and not glu_fused"; do
  run python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "$k"
done
run python -m pytest tests/test_gpu_attn_train.py -m gpu -q --tb=short -p no:cacheprovider
run python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider
run python bench.py --steps 5 --warmup 3 --skip-cpu
grep -E "^===|^exit|passed|failed|\"metric\"" gpurun_out/t6.log | cut -c1-1800
