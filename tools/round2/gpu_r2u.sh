#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sweep_r2.py --pairs 256 --batch 128 --reps 4 "default:" "c86:FFB_ITER_CARVEOUT=86" "c90:FFB_ITER_CARVEOUT=90" "c100:FFB_ITER_CARVEOUT=100" "default2:" "c86b:FFB_ITER_CARVEOUT=86" > gpurun_out/r2u_sweep_1080p.jsonl 2> gpurun_out/r2u_sweep_1080p.err
timeout 300 python tools/sweep_r2.py --size 640x360 --pairs 512 --batch 512 --reps 4 "default:" "c88:FFB_ITER_CARVEOUT=88" "c100:FFB_ITER_CARVEOUT=100" > gpurun_out/r2u_sweep_360p.jsonl 2> gpurun_out/r2u_sweep_360p.err
