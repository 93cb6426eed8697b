#!/bin/bash
# round 2, GPU call 2: wide-strip variants of k_flow_iter, per-level choices, segment / stream / batch settings
mkdir -p gpurun_out
timeout 900 python tools/sweep_r2.py --pairs 256 --reps 4 \
  "128x2x4:FFB_ITER_CFG=128x2x4" "256x4x8:FFB_ITER_CFG=256x4x8" "256x4x4:FFB_ITER_CFG=256x4x4" "256x2x4:FFB_ITER_CFG=256x2x4" \
  "512x4x8:FFB_ITER_CFG=512x4x8" "512x2x4:FFB_ITER_CFG=512x2x4" "192x4x8:FFB_ITER_CFG=192x4x8" \
  "256x4x8_seg2:FFB_ITER_CFG=256x4x8,FFB_ITER_MINSEG=2" "256x4x8_seg4:FFB_ITER_CFG=256x4x8,FFB_ITER_MINSEG=4" "256x4x8_seg5:FFB_ITER_CFG=256x4x8,FFB_ITER_MINSEG=5" \
  "256x4x8_seg6:FFB_ITER_CFG=256x4x8,FFB_ITER_MINSEG=6" "256x4x8_seg8:FFB_ITER_CFG=256x4x8,FFB_ITER_MINSEG=8" \
  "256x4x8_s1:FFB_ITER_CFG=256x4x8,FFB_FLOW_STREAMS=1" "256x4x8_s3:FFB_ITER_CFG=256x4x8,FFB_FLOW_STREAMS=3" "256x4x8_s4:FFB_ITER_CFG=256x4x8,FFB_FLOW_STREAMS=4" \
  "256x4x8_b128:FFB_ITER_CFG=256x4x8,BATCH=128" "256x4x8_b32:FFB_ITER_CFG=256x4x8,BATCH=32" "256x4x8_b128_s4:FFB_ITER_CFG=256x4x8,BATCH=128,FFB_FLOW_STREAMS=4" \
  "128x2x4_b128:FFB_ITER_CFG=128x2x4,BATCH=128" \
  "mix_k1_128:FFB_ITER_CFG=256x4x8,FFB_ITER_CFG_K1=128x2x4" "mix_k3_128:FFB_ITER_CFG=256x4x8,FFB_ITER_CFG_K3=128x2x4" \
  "mix_k23_128:FFB_ITER_CFG=256x4x8,FFB_ITER_CFG_K3=128x2x4,FFB_ITER_CFG_K2=128x2x4" "mix_k123_256x2x4:FFB_ITER_CFG=256x4x8,FFB_ITER_CFG_K1=256x2x4,FFB_ITER_CFG_K2=256x2x4,FFB_ITER_CFG_K3=256x2x4" \
  "mix_k3_96:FFB_ITER_CFG=256x4x8,FFB_ITER_CFG_K3=96x2x4" "mix_k0only:FFB_ITER_CFG=128x2x4,FFB_ITER_CFG_K0=256x4x8" \
  "256x4x8_opt:FFB_ITER_CFG=256x4x8,FFB_ITER_SWMAX=0" \
  > gpurun_out/r2b_sweep_1080p.jsonl 2> gpurun_out/r2b_sweep_1080p.err
timeout 600 python tools/sweep_r2.py --pairs 64 --batch 16 --reps 3 --size 3840x2160 \
  "128x2x4:FFB_ITER_CFG=128x2x4" "256x4x8:FFB_ITER_CFG=256x4x8" "256x2x4:FFB_ITER_CFG=256x2x4" "512x4x8:FFB_ITER_CFG=512x4x8" \
  > gpurun_out/r2b_sweep_4k.jsonl 2> gpurun_out/r2b_sweep_4k.err
timeout 600 python tools/sweep_r2.py --pairs 512 --batch 128 --reps 4 --size 640x360 \
  "160x2x4:FFB_ITER_CFG=160x2x4" "128x2x4:FFB_ITER_CFG=128x2x4" "256x4x8:FFB_ITER_CFG=256x4x8" "256x2x4:FFB_ITER_CFG=256x2x4" "192x4x8:FFB_ITER_CFG=192x4x8" \
  > gpurun_out/r2b_sweep_360p.jsonl 2> gpurun_out/r2b_sweep_360p.err
timeout 600 python tools/sweep_r2.py --size 256x256 --pairs 1024 --batch 512 --reps 4 \
  "160x2x4:FFB_ITER_CFG=160x2x4" "256x4x8:FFB_ITER_CFG=256x4x8" "192x4x8:FFB_ITER_CFG=192x4x8" "160x2x4_seg1:FFB_ITER_CFG=160x2x4,FFB_ITER_MINSEG=1" \
  "mix:FFB_ITER_CFG=160x2x4,FFB_ITER_CFG_K0=256x4x8" \
  > gpurun_out/r2b_sweep_256.jsonl 2> gpurun_out/r2b_sweep_256.err
echo done
