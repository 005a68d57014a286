#!/bin/bash
# One parameterised GPU trip (replaces the per-trip scripts of round 1).  Run it through gpurun:
#   gpurun --timeout 900 -- 'bash scripts/gpu_trip.sh <label> [tests] [bench] [bench8] [launches] [ncu:<kernel regex>] [memcheck] [racecheck] [py:<script> ...]'
# Every step writes gpurun_out/<label>_<step>.log; a step only runs when named.  ncu / compute-sanitizer steps run after
# the plain run of the same command has exited 0 (B200_PROFILING.md), one tool per trip.
set -u
label=${1:?label}; shift
mkdir -p gpurun_out
out=gpurun_out/${label}
for step in "$@"; do
  case "$step" in
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > ${out}_tests.log 2>&1; echo "tests exit $?"; tail -5 ${out}_tests.log ;;
    testsall)
      timeout 1200 python -m pytest tests -m gpu -q > ${out}_tests.log 2>&1; echo "tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" ${out}_tests.log | tail -40 ;;
    tests:*)
      timeout 900 python -m pytest tests -m gpu -x -q -k "${step#tests:}" > ${out}_tests.log 2>&1; echo "tests exit $?"; tail -15 ${out}_tests.log ;;
    bench)
      timeout 600 python bench.py --steps 10 --warmup 3 > ${out}_bench.log 2>&1; echo "bench exit $?"; tail -1 ${out}_bench.log | cut -c1-1500 ;;
    benchfast)
      timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e --skip-extras > ${out}_benchfast.log 2>&1; echo "bench exit $?"
      tail -1 ${out}_benchfast.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), d['phases_ms'], d['losses'])" ;;
    bench:*)
      timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-e2e --skip-extras --config "${step#bench:}" > ${out}_bench_${step#bench:}.log 2>&1; echo "bench exit $?"
      tail -1 ${out}_bench_${step#bench:}.log | cut -c1-600 ;;
    launches)
      timeout 600 python bench.py --ncu --skip-cpu --skip-e2e > ${out}_plain.log 2>&1 && \
      timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file ${out}_launches.csv python bench.py --ncu --skip-cpu --skip-e2e > ${out}_ncu.log 2>&1; echo "launches exit $?" ;;
    minibatch)
      timeout 600 python bench.py --ncu-minibatch > ${out}_plain_mb.log 2>&1 && \
      timeout 1500 ncu --profile-from-start off --clock-control none --csv --metrics \
gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
        --log-file ${out}_minibatch.csv python bench.py --ncu-minibatch > ${out}_ncu_mb.log 2>&1; echo "minibatch exit $?" ;;
    decodestep)
      timeout 600 python bench.py --ncu-decode-step > ${out}_plain_ds.log 2>&1 && \
      timeout 1500 ncu --profile-from-start off --clock-control none --csv --metrics \
gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
        --log-file ${out}_decodestep.csv python bench.py --ncu-decode-step > ${out}_ncu_ds.log 2>&1; echo "decodestep exit $?" ;;
    ncu:*)
      timeout 600 python bench.py --ncu-minibatch > ${out}_plain_full.log 2>&1 && \
      timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:${step#ncu:}" -c 2 \
        -o ${out}_full python bench.py --ncu-minibatch > ${out}_ncu_full.log 2>&1; echo "ncu exit $?" ;;
    memcheck|racecheck|synccheck|initcheck)
      timeout 1700 compute-sanitizer --tool $step --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "${SANITIZE_K:-fused or trainer}" \
        > ${out}_${step}.log 2>&1; echo "$step exit $?"; tail -8 ${out}_${step}.log ;;
    py:*)
      s=${step#py:}; timeout 900 python $s > ${out}_$(basename ${s%% *} .py).log 2>&1; echo "$s exit $?"; tail -20 ${out}_$(basename ${s%% *} .py).log ;;
    *) echo "unknown step $step" ;;
  esac
done
