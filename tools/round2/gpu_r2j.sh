#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/sweep_r2.py --size 256x256 --pairs 1024 --batch 512 --reps 5 \
  "default:" "k3_64:FFB_ITER_CFG_K3=64x2x4" "k3_64_k2_96:FFB_ITER_CFG_K3=64x2x4,FFB_ITER_CFG_K2=96x2x4" "k2_96:FFB_ITER_CFG_K2=96x2x4" \
  "k3_64_k2_96_seg1:FFB_ITER_CFG_K3=64x2x4,FFB_ITER_CFG_K2=96x2x4,FFB_ITER_MINSEG=1" "k23_64:FFB_ITER_CFG_K3=64x2x4,FFB_ITER_CFG_K2=64x2x4" \
  "k3_64_k2_96_k1_160_s1:FFB_ITER_CFG_K3=64x2x4,FFB_ITER_CFG_K2=96x2x4,FFB_FLOW_STREAMS=1" "all_seg1_s1:FFB_ITER_CFG_K3=64x2x4,FFB_ITER_CFG_K2=96x2x4,FFB_ITER_MINSEG=1,FFB_FLOW_STREAMS=1" \
  "best_b1024:FFB_ITER_CFG_K3=64x2x4,FFB_ITER_CFG_K2=96x2x4,FFB_ITER_MINSEG=1,BATCH=1024" \
  > gpurun_out/r2j_sweep_256.jsonl 2> gpurun_out/r2j_sweep_256.err
timeout 400 python tools/sweep_r2.py --size 640x360 --pairs 512 --batch 256 --reps 4 \
  "default:" "k3_96:FFB_ITER_CFG_K3=96x2x4" "k3_96_k2_160:FFB_ITER_CFG_K3=96x2x4,FFB_ITER_CFG_K2=160x2x4" "k3_96_seg1:FFB_ITER_CFG_K3=96x2x4,FFB_ITER_MINSEG=1" \
  "s1:FFB_FLOW_STREAMS=1" "k3_96_s1:FFB_ITER_CFG_K3=96x2x4,FFB_FLOW_STREAMS=1" "b512:BATCH=512" "k0_256:FFB_ITER_CFG_K0=256x4x8" \
  > gpurun_out/r2j_sweep_360p.jsonl 2> gpurun_out/r2j_sweep_360p.err
timeout 300 python tools/sweep_r2.py --pairs 256 --batch 128 --reps 4 \
  "default:" "k3_96:FFB_ITER_CFG_K3=96x2x4" "k3_128:FFB_ITER_CFG_K3=128x2x4" "k3_160:FFB_ITER_CFG_K3=160x2x4" \
  > gpurun_out/r2j_sweep_1080p.jsonl 2> gpurun_out/r2j_sweep_1080p.err
echo done
