#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2s_smoke.log
timeout 400 python tools/sweep_r2.py --pairs 128 --batch 64 --reps 3 --size 3840x2160 "default:" > gpurun_out/r2s_sweep_4k.jsonl 2> gpurun_out/r2s_sweep_4k.err
timeout 400 python tools/sweep_r2.py --pairs 72 --batch 36 --reps 3 --size 5760x2880 "default:" > gpurun_out/r2s_sweep_5760.jsonl 2> gpurun_out/r2s_sweep_5760.err
cat gpurun_out/r2s_sweep_4k.jsonl gpurun_out/r2s_sweep_5760.jsonl | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s_pytest.log
