#!/bin/bash
# 8-GPU call: strong scaling of one bracket (c2-strong), whole-library throughput (c5), multi-rank tests over NCCL
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2m_gpus.txt; nproc >> gpurun_out/r2m_gpus.txt
TR() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for n in 8 4 2; do
  timeout 400 bash -c "$(declare -f TR); TR $n 2970$n bench.py --gpus $n --workload c2-strong --strong-pairs 1024 --steps 4 --warmup 1" > gpurun_out/r2m_strong_n$n.json 2> gpurun_out/r2m_strong_n$n.err; echo "strong n=$n rc=$?"
done
timeout 400 python bench.py --workload c2-strong --strong-pairs 1024 --steps 4 --warmup 1 > gpurun_out/r2m_strong_n1.json 2> gpurun_out/r2m_strong_n1.err; echo "strong n=1 rc=$?"
timeout 400 bash -c "$(declare -f TR); TR 8 29718 bench.py --gpus 8 --workload c5 --steps 3" > gpurun_out/r2m_c5_n8.json 2> gpurun_out/r2m_c5_n8.err; echo "c5 n=8 rc=$?"
timeout 400 python bench.py --workload c5 --steps 2 > gpurun_out/r2m_c5_n1.json 2> gpurun_out/r2m_c5_n1.err; echo "c5 n=1 rc=$?"
timeout 400 bash -c "$(declare -f TR); TR 8 29728 bench.py --gpus 8 --steps 10 --warmup 3" > gpurun_out/r2m_c2_n8.json 2> gpurun_out/r2m_c2_n8.err; echo "c2 n=8 rc=$?"
timeout 300 python bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2m_ref.json 2> gpurun_out/r2m_ref.err; echo "ref rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "torchrun or sharded_over_ranks or frame_range_shards" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest.log
