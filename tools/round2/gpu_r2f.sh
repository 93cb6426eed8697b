#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2f_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"
for t in 1 2 8; do FFB_COPY_THREADS=$t timeout 600 python bench.py --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/r2f_bench_ct$t.json 2> gpurun_out/r2f_bench_ct$t.err; done
timeout 600 python bench.py --workload c5 --c5-videos 16 --steps 2 > gpurun_out/r2f_bench_c5.json 2> gpurun_out/r2f_bench_c5.err
echo "c5 rc=$?"
