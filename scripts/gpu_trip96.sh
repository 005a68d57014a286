#!/bin/bash
# trip 96: policy-head kernel with the max-shared carve-out (same L1 / shared split as the rest of the step); PDL on it again?
mkdir -p gpurun_out
run() { # carveout pdl
  MPX_HEAD_CARVEOUT=$1 MPX_PDL=$2 timeout 300 python bench.py --skip-cpu --skip-e2e --steps 10 --warmup 3 > gpurun_out/t96_c$1_p$2.log 2>&1
  echo -n "carveout $1 pdl $2: "; tail -1 gpurun_out/t96_c$1_p$2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'])"
}
run 0 12
run 1 12
run 1 28
run 1 14
run 0 12
run 1 12
