#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attn_decode_fused" > gpurun_out/t59_tests.log 2>&1; echo "tests exit $?"; tail -15 gpurun_out/t59_tests.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_golden.py tests/test_gpu_trainer.py -m gpu -q -x > gpurun_out/t59_tests2.log 2>&1; echo "tests2 exit $?"; tail -4 gpurun_out/t59_tests2.log
timeout 600 python bench.py > gpurun_out/t59_bench.log 2>&1; tail -1 gpurun_out/t59_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['phases_ms'], d['roofline']['frac'], d['roofline']['avg_launch_us'])"
MPX_PDL=1 timeout 600 python bench.py --skip-cpu > gpurun_out/t59_bench_pdl.log 2>&1; tail -1 gpurun_out/t59_bench_pdl.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PDL', d['value'], d['e2e']['value'], d['phases_ms'])"
