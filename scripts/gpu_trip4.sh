#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/probe.log gpurun_out/t4.log
run() { echo "=== $*" >> gpurun_out/t4.log; timeout 900 "$@" >> gpurun_out/t4.log 2>&1; echo "exit $?" >> gpurun_out/t4.log; }
run python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
run python bench.py --steps 5 --warmup 3 --skip-cpu
python bench.py --ncu > gpurun_out/plain_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 14226 -c 14400 --csv --log-file gpurun_out/launches2.csv python bench.py --ncu > gpurun_out/ncu1.log 2>&1
echo "ncu1 exit $?" >> gpurun_out/t4.log
python bench.py --ncu > gpurun_out/plain_ncu2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_decode -s 1600 -c 2 -o gpurun_out/prof_attn_decode_mma python bench.py --ncu > gpurun_out/ncu2.log 2>&1
echo "ncu2 exit $?" >> gpurun_out/t4.log
grep -E "^===|^exit|passed|failed|ncu|\"metric\"" gpurun_out/t4.log | cut -c1-1500
