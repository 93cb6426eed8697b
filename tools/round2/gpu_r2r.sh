#!/bin/bash
# final 8-GPU evidence of the round: weak c2, whole-library c5, one 3000-pair bracket over 8 GPUs, reference arm on the same box
mkdir -p gpurun_out
nproc > gpurun_out/r2r_host.txt; nvidia-smi -L >> gpurun_out/r2r_host.txt
TR() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
timeout 400 bash -c "$(declare -f TR); TR 8 29728 bench.py --gpus 8 --steps 12 --warmup 4" > gpurun_out/r2r_c2_n8.json 2> gpurun_out/r2r_c2_n8.err; echo "c2 n=8 rc=$?"
timeout 300 python bench.py --steps 12 --warmup 4 --no-extras --no-cpu-baseline > gpurun_out/r2r_c2_n1.json 2> gpurun_out/r2r_c2_n1.err; echo "c2 n=1 rc=$?"
timeout 400 bash -c "$(declare -f TR); TR 8 29718 bench.py --gpus 8 --workload c5 --steps 3" > gpurun_out/r2r_c5_n8.json 2> gpurun_out/r2r_c5_n8.err; echo "c5 n=8 rc=$?"
timeout 400 python bench.py --workload c5 --steps 2 > gpurun_out/r2r_c5_n1.json 2> gpurun_out/r2r_c5_n1.err; echo "c5 n=1 rc=$?"
timeout 400 bash -c "$(declare -f TR); TR 8 29738 bench.py --gpus 8 --workload c2-strong --strong-pairs 3000 --steps 3 --warmup 1" > gpurun_out/r2r_strong3000_n8.json 2> gpurun_out/r2r_strong3000_n8.err; echo "strong n=8 rc=$?"
timeout 300 python bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2r_ref.json 2> gpurun_out/r2r_ref.err; echo "ref rc=$?"
