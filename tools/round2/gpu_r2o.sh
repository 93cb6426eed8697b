#!/bin/bash
mkdir -p gpurun_out
TR() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
timeout 400 bash -c "$(declare -f TR); TR 8 29728 bench.py --gpus 8 --steps 12 --warmup 3" > gpurun_out/r2o_c2_n8.json 2> gpurun_out/r2o_c2_n8.err; echo "c2 n=8 rc=$?"
timeout 400 bash -c "$(declare -f TR); TR 8 29718 bench.py --gpus 8 --workload c5 --steps 3" > gpurun_out/r2o_c5_n8.json 2> gpurun_out/r2o_c5_n8.err; echo "c5 n=8 rc=$?"
timeout 400 python bench.py --workload c5 --steps 2 > gpurun_out/r2o_c5_n1.json 2> gpurun_out/r2o_c5_n1.err; echo "c5 n=1 rc=$?"
timeout 400 bash -c "$(declare -f TR); TR 8 29738 bench.py --gpus 8 --workload c2-strong --strong-pairs 3000 --steps 3 --warmup 1" > gpurun_out/r2o_strong3000_n8.json 2> gpurun_out/r2o_strong3000_n8.err; echo "strong n=8 rc=$?"
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2o_c2_n1.json 2> gpurun_out/r2o_c2_n1.err; echo "c2 n=1 rc=$?"
