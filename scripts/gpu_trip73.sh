#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attn_train.py tests/test_gpu_model.py tests/test_gpu_golden.py -m gpu -q -x > gpurun_out/t73_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t73_tests.log
timeout 300 python scripts/stamps_decode.py 2>&1 | grep -A32 "attn_bwd_tc" | awk 'NR%5==2 || NR<3' | head -12
timeout 600 python bench.py --skip-cpu > gpurun_out/t73_bench.log 2>&1; tail -1 gpurun_out/t73_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['phases_ms'])"
