#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/t8.log gpurun_out/probe.log
run() { echo "=== $*" >> gpurun_out/t8.log; timeout 420 "$@" >> gpurun_out/t8.log 2>&1; echo "exit $?" >> gpurun_out/t8.log; }
run python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider -k "conv1 or train_forward or decode_rollout"
python bench.py --ncu-minibatch > gpurun_out/plain_mb.log 2>&1 && \
ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof_minibatch python bench.py --ncu-minibatch > gpurun_out/ncu_mb.log 2>&1
echo "ncu mb exit $?" >> gpurun_out/t8.log
python bench.py --ncu-decode-step > gpurun_out/plain_ds.log 2>&1 && \
ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof_decode_step python bench.py --ncu-decode-step > gpurun_out/ncu_ds.log 2>&1
echo "ncu ds exit $?" >> gpurun_out/t8.log
grep -E "^===|^exit|passed|failed|ncu" gpurun_out/t8.log | cut -c1-300
ls -la gpurun_out | head -20
