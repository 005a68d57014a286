#!/bin/bash
# trip 84: programmatic dependent launch also on the GEMM / elementwise / value kernels of the update graph (bit 1)?
mkdir -p gpurun_out
for m in 12 13 12 13; do
  MPX_PDL=$m timeout 300 python bench.py --skip-cpu --skip-e2e --steps 10 --warmup 3 > gpurun_out/t84_$m.log 2>&1
  echo -n "PDL mask $m: "; tail -1 gpurun_out/t84_$m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'], d['losses'])"
done
MPX_PDL=13 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t84_tests13.log 2>&1; echo "tests (mask 13) exit $?"; tail -2 gpurun_out/t84_tests13.log
