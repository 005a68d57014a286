#!/bin/bash
# trip 79: encoder weight gradients + policy-head backward on side streams; A/B of the encoder part, 10 timed steps each
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t79_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t79_tests.log
for m in 0 1 0 1; do
  MPX_ENC_WGRAD_OVERLAP=$m timeout 300 python bench.py --skip-cpu --skip-e2e --steps 10 --warmup 3 > gpurun_out/t79_enc$m.log 2>&1
  echo -n "encoder wgrad overlap $m: "; tail -1 gpurun_out/t79_enc$m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'])"
done
