#!/bin/bash
# trip 83: decode attention starts its cache stream before griddepcontrol.wait (flags bit 0); A/B with MPX_EARLY_KV
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t83_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t83_tests.log
for m in 0 1 0 1; do
  MPX_EARLY_KV=$m timeout 300 python bench.py --skip-cpu --skip-e2e --steps 10 --warmup 3 > gpurun_out/t83_ek$m.log 2>&1
  echo -n "early KV stream $m: "; tail -1 gpurun_out/t83_ek$m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'], round(d['roofline']['frac'],4), d['losses'])"
done
