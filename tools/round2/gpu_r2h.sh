#!/bin/bash
mkdir -p gpurun_out
# the flag-synchronised variant first on a small case under a short timeout (a deadlock must not hang the box)
FFB_ITER_FLAGS=1 timeout 120 python tools/sweep_r2.py --pairs 8 --batch 8 --reps 1 "flags_small:FFB_ITER_FLAGS=1" > gpurun_out/r2h_small.jsonl 2> gpurun_out/r2h_small.err
echo "small rc=$?"; cat gpurun_out/r2h_small.jsonl
timeout 300 python tools/sweep_r2.py --pairs 256 --batch 128 --reps 5 \
  "barriers:" "flags:FFB_ITER_FLAGS=1" "barriers2:" "flags2:FFB_ITER_FLAGS=1" "flags_s1:FFB_ITER_FLAGS=1,FFB_FLOW_STREAMS=1" "barriers_s1:FFB_FLOW_STREAMS=1" \
  > gpurun_out/r2h_sweep_1080p.jsonl 2> gpurun_out/r2h_sweep_1080p.err
echo "sweep rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "torchrun or sharded_over_ranks or frame_range_shards or oracle_1080p" > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2h_pytest.log
