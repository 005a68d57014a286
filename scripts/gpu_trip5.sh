#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/t5.log
run() { echo "=== $*" >> gpurun_out/t5.log; timeout 900 "$@" >> gpurun_out/t5.log 2>&1; echo "exit $?" >> gpurun_out/t5.log; }
nvidia-smi --query-gpu=index,name --format=csv >> gpurun_out/t5.log 2>&1
run python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider -x
run python bench.py --steps 5 --warmup 3 --skip-cpu --skip-e2e
run python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu
grep -E "^===|^exit|passed|failed|\"metric\"" gpurun_out/t5.log | cut -c1-900
