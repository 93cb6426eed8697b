#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_video.py > gpurun_out/r2w_video.json 2> gpurun_out/r2w_video.err; echo "video rc=$?"; cat gpurun_out/r2w_video.json
timeout 600 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2w_ref.json 2> gpurun_out/r2w_ref.err; echo "ref rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2w_pytest.log
