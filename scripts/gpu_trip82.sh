#!/bin/bash
# trip 82: GPU suite + bench with programmatic dependent launch on the decode attention / fused GLU kernels by default
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t82_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t82_tests.log
timeout 300 python bench.py --skip-cpu --steps 10 --warmup 3 > gpurun_out/t82_bench.log 2>&1
echo -n "bench: "; tail -1 gpurun_out/t82_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), round(d['e2e']['value']/1e6,3), d['phases_ms'], d['roofline']['frac'])"
timeout 300 python bench.py --config return_8_layers --skip-cpu --skip-e2e --steps 4 --warmup 3 > gpurun_out/t82_r8.log 2>&1
echo -n "return_8: "; tail -1 gpurun_out/t82_r8.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'])"
MPX_PDL=0 timeout 300 python bench.py --config return_8_layers --skip-cpu --skip-e2e --steps 4 --warmup 3 > gpurun_out/t82_r8_pdl0.log 2>&1
echo -n "return_8 PDL 0: "; tail -1 gpurun_out/t82_r8_pdl0.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'])"
