#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sweep_r2.py --pairs 256 --reps 5 \
  "poly_scalar:FFB_POLY=0" "poly_packed:" "poly_scalar2:FFB_POLY=0" "poly_packed2:" "packed_b128:BATCH=128" \
  > gpurun_out/r2e_sweep_1080p.jsonl 2> gpurun_out/r2e_sweep_1080p.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err
echo "ref rc=$?"
timeout 600 python bench.py --workload c2-strong --strong-pairs 512 --steps 3 --warmup 1 > gpurun_out/r2e_bench_strong.json 2> gpurun_out/r2e_bench_strong.err
echo "strong rc=$?"
timeout 600 python bench.py --workload c5 --c5-videos 16 --steps 2 > gpurun_out/r2e_bench_c5.json 2> gpurun_out/r2e_bench_c5.err
echo "c5 rc=$?"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest.log
