#!/bin/bash
# round 2, GPU call 1: k_flow_iter variants at 1080p and 256x256, then the GPU test-suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 900 python tools/sweep_r2.py --pairs 128 --reps 5 \
  "r1rule:FFB_ITER_SWMAX=0" "default:" "default_again:" "opt1:FFB_ITER_OPT=1" "opt2:FFB_ITER_OPT=2" "opt3:FFB_ITER_OPT=3" \
  "128x2x8:FFB_ITER_CFG=128x2x8" "128x4x8:FFB_ITER_CFG=128x4x8" "128x4x4:FFB_ITER_CFG=128x4x4" \
  "256x2x4:FFB_ITER_CFG=256x2x4" "256x2x8:FFB_ITER_CFG=256x2x8" "256x4x8:FFB_ITER_CFG=256x4x8" \
  "160x2x4:FFB_ITER_CFG=160x2x4" "160x2x8:FFB_ITER_CFG=160x2x8" \
  "coarse256x2x8:FFB_ITER_CFG_COARSE=256x2x8" "coarse256x2x4:FFB_ITER_CFG_COARSE=256x2x4" "coarse160x2x4:FFB_ITER_CFG_COARSE=160x2x4" \
  "coarse96x2x4:FFB_ITER_CFG_COARSE=96x2x4" \
  "sh540:FFB_ITER_SH=540,FFB_ITER_MINSEG=2" "seg4:FFB_ITER_MINSEG=4" \
  "s1_default:FFB_FLOW_STREAMS=1" "s1_r1rule:FFB_FLOW_STREAMS=1,FFB_ITER_SWMAX=0" "s1_128x4x8:FFB_FLOW_STREAMS=1,FFB_ITER_CFG=128x4x8" \
  "s1_256x2x8:FFB_FLOW_STREAMS=1,FFB_ITER_CFG=256x2x8" "s1_128x2x8:FFB_FLOW_STREAMS=1,FFB_ITER_CFG=128x2x8" "s1_opt3:FFB_FLOW_STREAMS=1,FFB_ITER_OPT=3" \
  > gpurun_out/r2a_sweep_1080p.jsonl 2> gpurun_out/r2a_sweep_1080p.err
timeout 600 python tools/sweep_r2.py --size 256x256 --pairs 1024 --batch 512 --reps 5 \
  "r1rule:FFB_ITER_SWMAX=0" "default:" "128x2x4:FFB_ITER_CFG=128x2x4" "160x2x8:FFB_ITER_CFG=160x2x8" "256x2x4:FFB_ITER_CFG=256x2x4" "256x2x8:FFB_ITER_CFG=256x2x8" \
  "96x2x4:FFB_ITER_CFG=96x2x4" "opt3_128:FFB_ITER_CFG=128x2x4,FFB_ITER_OPT=3" \
  > gpurun_out/r2a_sweep_256.jsonl 2> gpurun_out/r2a_sweep_256.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -3 gpurun_out/r2a_pytest.log
