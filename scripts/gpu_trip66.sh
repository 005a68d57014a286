#!/bin/bash
# trip 66: shared-memory bank conflict census of the decode step and the minibatch
mkdir -p gpurun_out
M="gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active"
python bench.py --ncu-decode-step > gpurun_out/plain_ds.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/ds_conf_t66.csv python bench.py --ncu-decode-step > gpurun_out/ncu_ds.log 2>&1
echo "ncu ds exit $?"
python bench.py --ncu-minibatch > gpurun_out/plain_mb.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/mb_conf_t66.csv python bench.py --ncu-minibatch > gpurun_out/ncu_mb.log 2>&1
echo "ncu mb exit $?"
