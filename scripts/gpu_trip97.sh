#!/bin/bash
# trip 97: final tree: GPU suite, smoke, default bench line; rollout chains knob once more (now with PDL)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t97_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/t97_tests.log
timeout 300 python scripts/run_smoke.py > gpurun_out/t97_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/t97_smoke.log
timeout 600 python bench.py > gpurun_out/t97_bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/t97_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), round(d['e2e']['value']/1e6,3), d['phases_ms'], d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'])"
MPX_ROLLOUT_CHAINS=2 timeout 300 python bench.py --skip-cpu --skip-e2e --steps 10 --warmup 3 > gpurun_out/t97_chains2.log 2>&1
echo -n "2 rollout chains: "; tail -1 gpurun_out/t97_chains2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'])"
