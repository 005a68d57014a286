#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/t7.log gpurun_out/probe.log
run() { echo "=== $*" >> gpurun_out/t7.log; timeout 420 "$@" >> gpurun_out/t7.log 2>&1; echo "exit $?" >> gpurun_out/t7.log; }
run python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider
run python bench.py --steps 5 --warmup 3 --skip-cpu --skip-e2e
run python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu --skip-e2e
grep -E "^===|^exit|passed|failed|\"metric\"|Error" gpurun_out/t7.log | cut -c1-1700
