#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2p_bench.err
timeout 600 python bench.py --workload c5 --steps 2 > gpurun_out/r2p_c5.json 2> gpurun_out/r2p_c5.err; echo "c5 rc=$?"; tail -2 gpurun_out/r2p_c5.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2p_pytest.log
