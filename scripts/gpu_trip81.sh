#!/bin/bash
# trip 81: programmatic dependent launch revisited with 10 timed steps per run (trip 79 showed +-0.03 ms run-to-run noise on one box)
mkdir -p gpurun_out
for m in 0 4 12 0 4 12 8; do
  MPX_PDL=$m timeout 300 python bench.py --skip-cpu --skip-e2e --steps 10 --warmup 3 > gpurun_out/t81_$m.log 2>&1
  echo -n "PDL mask $m: "; tail -1 gpurun_out/t81_$m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'], round(d['roofline']['frac'],4))"
done
