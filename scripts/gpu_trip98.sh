#!/bin/bash
# trip 98: HL-Gauss loss kernel: one reciprocal per row, softmax from the exponentials already at hand, three reductions in one butterfly, packed bf16 gradient stores
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t98_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t98_tests.log
for i in 1 2; do
timeout 300 python bench.py --skip-cpu --skip-e2e --steps 10 --warmup 3 > gpurun_out/t98_bench$i.log 2>&1
echo -n "bench: "; tail -1 gpurun_out/t98_bench$i.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'], d['losses'])"
done
