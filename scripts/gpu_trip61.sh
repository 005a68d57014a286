#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/microbench_attn_decode.py > gpurun_out/t61_attn.log 2>&1; cat gpurun_out/t61_attn.log
timeout 300 python scripts/microbench_attn_decode.py --fused > gpurun_out/t61_attn_fused.log 2>&1; cat gpurun_out/t61_attn_fused.log
