#!/bin/bash
# trip 90: gelu transpose fused into the dgrad GEMM epilogue (1) vs linear + gelu_bwd (0); default bench line
mkdir -p gpurun_out
for m in 1 0 1 0; do
MPX_FUSE_GELU_T=$m timeout 300 python bench.py --skip-cpu --skip-e2e --steps 10 --warmup 3 > gpurun_out/t90_fg$m.log 2>&1
echo -n "fused gelu transpose $m: "; tail -1 gpurun_out/t90_fg$m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'])"
done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t90_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/t90_tests.log
time (timeout 900 python bench.py > gpurun_out/t90_bench.log 2> gpurun_out/t90_bench.err); echo "bench exit $?"; tail -1 gpurun_out/t90_bench.log | cut -c1-2600
