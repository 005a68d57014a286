#!/bin/bash
# trip 92: the other named configs with the end-of-round code (1 GPU each)
mkdir -p gpurun_out
for c in return_8_layers multitask blog_32_layers; do
  timeout 500 python bench.py --config $c --skip-cpu --skip-e2e --steps 3 --warmup 3 > gpurun_out/t92_$c.log 2>&1
  echo -n "$c exit $?: "; tail -1 gpurun_out/t92_$c.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'], d['config']['workload'][:80])"
done
