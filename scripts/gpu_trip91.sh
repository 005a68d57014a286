#!/bin/bash
# trip 91 (2 GPUs): end-of-round code under torchrun, as the driver launches it
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/t91_2gpu.log 2>&1
echo "2-GPU exit $?"; tail -1 gpurun_out/t91_2gpu.log | cut -c1-1800
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/t91_2gpu_ref.log 2>&1
echo "2-GPU reference arm exit $?"; tail -1 gpurun_out/t91_2gpu_ref.log | cut -c1-300
