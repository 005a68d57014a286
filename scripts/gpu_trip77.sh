#!/bin/bash
# trip 77: GPU suite after the phased Muon step / gelu-transpose GEMM epilogue, then A/B of the Muon overlap
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t77_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t77_tests.log
for m in 0 1 0 1; do
  MPX_MUON_OVERLAP=$m timeout 300 python bench.py --skip-cpu --skip-e2e --steps 5 --warmup 3 > gpurun_out/t77_ov$m.log 2>&1
  echo -n "muon overlap $m: "; tail -1 gpurun_out/t77_ov$m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'])"
done
