#!/bin/bash
# trip 88: capture the graphs on a high-priority stream (main chain outranks the side-stream weight gradients)?
mkdir -p gpurun_out
for m in 0 1 0 1; do
MPX_HP_CAPTURE=$m timeout 300 python bench.py --skip-cpu --skip-e2e --steps 10 --warmup 3 > gpurun_out/t88_hp$m.log 2>&1
echo -n "hp capture $m: "; tail -1 gpurun_out/t88_hp$m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'])"
done
