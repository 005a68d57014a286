#!/bin/bash
# trip 76: which kernel families gain from programmatic dependent launch in the rollout graph
mkdir -p gpurun_out
for m in 0 2 4 8 16 30; do
  MPX_PDL=$m timeout 300 python bench.py --skip-cpu --skip-e2e --steps 3 --warmup 3 > gpurun_out/t76_$m.log 2>&1
  echo -n "PDL mask $m: "; tail -1 gpurun_out/t76_$m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'])"
done
