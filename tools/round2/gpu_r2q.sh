#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sweep_r2.py --pairs 512 --batch 128 --reps 4 \
  "b128:" "b192:BATCH=192" "b256:BATCH=256" "b256_s4:BATCH=256,FFB_FLOW_STREAMS=4" "b128_s3:FFB_FLOW_STREAMS=3" "b128_again:" "b256_s3:BATCH=256,FFB_FLOW_STREAMS=3" \
  > gpurun_out/r2q_sweep_1080p.jsonl 2> gpurun_out/r2q_sweep_1080p.err
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv > gpurun_out/r2q_clocks.txt
