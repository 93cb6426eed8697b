#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sweep_r2.py --pairs 256 --batch 128 --reps 6 \
  "radial_side:" "radial_inline:FFB_RADIAL_STREAM=0" "radial_side2:" "radial_inline2:FFB_RADIAL_STREAM=0" "s1:FFB_FLOW_STREAMS=1" \
  > gpurun_out/r2i_sweep_1080p.jsonl 2> gpurun_out/r2i_sweep_1080p.err
timeout 300 python tools/sweep_r2.py --size 256x256 --pairs 1024 --batch 512 --reps 5 \
  "default:" "inline:FFB_RADIAL_STREAM=0" "s1:FFB_FLOW_STREAMS=1" "s1_128:FFB_FLOW_STREAMS=1,FFB_ITER_CFG=128x2x4" "seg1:FFB_ITER_MINSEG=1" "b1024:BATCH=1024" "s4:FFB_FLOW_STREAMS=4" \
  > gpurun_out/r2i_sweep_256.jsonl 2> gpurun_out/r2i_sweep_256.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2i_pytest.log
