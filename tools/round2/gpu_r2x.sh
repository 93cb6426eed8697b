#!/bin/bash
mkdir -p gpurun_out
TR() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
timeout 300 python bench.py --steps 12 --warmup 4 --no-extras --no-cpu-baseline > gpurun_out/r2x_c2_n1.json 2> gpurun_out/r2x_c2_n1.err; echo "n1 rc=$?"
timeout 300 bash -c "$(declare -f TR); TR 2 29802 bench.py --gpus 2 --steps 12 --warmup 4 --no-extras" > gpurun_out/r2x_c2_n2.json 2> gpurun_out/r2x_c2_n2.err; echo "n2 rc=$?"
timeout 300 bash -c "$(declare -f TR); TR 4 29804 bench.py --gpus 4 --steps 12 --warmup 4 --no-extras" > gpurun_out/r2x_c2_n4.json 2> gpurun_out/r2x_c2_n4.err; echo "n4 rc=$?"
