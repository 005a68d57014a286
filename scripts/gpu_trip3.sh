#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/probe.log gpurun_out/t3.log
run() { echo "=== $*" >> gpurun_out/t3.log; timeout 900 "$@" >> gpurun_out/t3.log 2>&1; echo "exit $?" >> gpurun_out/t3.log; }
run python -m pytest tests/test_gpu_attn_train.py tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider
# launch list of ONE eager PPO update (second pass of the --ncu run; the first 15252 launches are the warm pass)
python bench.py --ncu > gpurun_out/plain_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 15252 -c 15400 --csv --log-file gpurun_out/launches.csv python bench.py --ncu > gpurun_out/ncu1.log 2>&1
echo "ncu1 exit $?" >> gpurun_out/t3.log
python bench.py --ncu > gpurun_out/plain_ncu2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_decode -s 1600 -c 3 -o gpurun_out/prof_attn_decode python bench.py --ncu > gpurun_out/ncu2.log 2>&1
echo "ncu2 exit $?" >> gpurun_out/t3.log
grep -E "^===|^exit|passed|failed|ncu" gpurun_out/t3.log | cut -c1-300
ls -la gpurun_out
