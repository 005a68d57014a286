#!/bin/bash
# trip 78 (2 GPUs): split gradient all-reduce (Muon partition early on the side stream) captured in the update graph
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "muon or gelu_transpose or optimizer" > gpurun_out/t78_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/t78_tests.log
for m in 1 0; do
  MPX_MUON_OVERLAP=$m timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --skip-cpu --steps 6 --warmup 3 > gpurun_out/t78_2gpu_ov$m.log 2>&1
  echo "2-GPU overlap $m exit $?"; tail -1 gpurun_out/t78_2gpu_ov$m.log | cut -c1-400
done
