#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --batch-frames 64 > gpurun_out/r2n_bench_b64.json 2> gpurun_out/r2n_bench_b64.err; echo "bench64 rc=$?"
timeout 600 python bench.py --workload c5 --steps 2 > gpurun_out/r2n_c5_pinned.json 2> gpurun_out/r2n_c5_pinned.err; echo "c5 rc=$?"
timeout 600 python bench.py --workload c5 --steps 2 --c5-pageable > gpurun_out/r2n_c5_pageable.json 2> gpurun_out/r2n_c5_pageable.err; echo "c5p rc=$?"
