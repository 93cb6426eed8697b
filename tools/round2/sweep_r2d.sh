#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/sweep_r2.py --pairs 256 --reps 5 \
  "256x4x8:FFB_ITER_CFG=256x4x8" "opt2:FFB_ITER_CFG=256x4x8,FFB_ITER_OPT=2" "opt4:FFB_ITER_CFG=256x4x8,FFB_ITER_OPT=4" "opt6:FFB_ITER_CFG=256x4x8,FFB_ITER_OPT=6" \
  "256x4x8_again:FFB_ITER_CFG=256x4x8" "b128_s4:BATCH=128,FFB_FLOW_STREAMS=4" "b128_s3:BATCH=128,FFB_FLOW_STREAMS=3" "b128_s2:BATCH=128" "b96_s3:BATCH=96,FFB_FLOW_STREAMS=3" \
  "b128_s4_opt4:BATCH=128,FFB_FLOW_STREAMS=4,FFB_ITER_OPT=4" "b128_s4_opt2:BATCH=128,FFB_FLOW_STREAMS=4,FFB_ITER_OPT=2" \
  > gpurun_out/r2d_sweep_1080p.jsonl 2> gpurun_out/r2d_sweep_1080p.err
timeout 600 python tools/sweep_r2.py --pairs 384 --batch 128 --reps 4 --size 1280x720 \
  "default:" "128x2x4:FFB_ITER_CFG=128x2x4" "256x4x8:FFB_ITER_CFG=256x4x8" "160x2x4:FFB_ITER_CFG=160x2x4" "256x2x4:FFB_ITER_CFG=256x2x4" \
  > gpurun_out/r2d_sweep_720p.jsonl 2> gpurun_out/r2d_sweep_720p.err
echo done
