#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/probe.log gpurun_out/t2.log
run() { echo "=== $*" >> gpurun_out/t2.log; timeout 600 "$@" >> gpurun_out/t2.log 2>&1; echo "exit $?" >> gpurun_out/t2.log; }
run python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider
run python -m pytest tests/test_gpu_attn_train.py -m gpu -q --tb=short -p no:cacheprovider
run python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider
run python bench.py --config test --steps 2 --warmup 1 --skip-cpu --no-graphs
run python bench.py --config test --steps 2 --warmup 1 --skip-cpu
run python bench.py --steps 2 --warmup 3 --skip-cpu --no-graphs --skip-e2e
run python bench.py --steps 3 --warmup 3 --skip-cpu
grep -E "^===|^exit|passed|failed|\"metric\"" gpurun_out/t2.log | cut -c1-400
