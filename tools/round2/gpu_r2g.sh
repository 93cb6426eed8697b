#!/bin/bash
# 2-GPU call: multi-rank tests with real NCCL, bench workloads at N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2g_gpus.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "torchrun or sharded_over_ranks or frame_range_shards" > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2g_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601"
timeout 600 $TR bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "c2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload c2-strong --strong-pairs 768 --steps 3 --warmup 1 > gpurun_out/r2g_strong_n2.json 2> gpurun_out/r2g_strong_n2.err; echo "strong rc=$?"
timeout 600 python bench.py --workload c2-strong --strong-pairs 768 --steps 3 --warmup 1 > gpurun_out/r2g_strong_n1.json 2> gpurun_out/r2g_strong_n1.err; echo "strong1 rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload c5 --c5-videos 32 --steps 2 > gpurun_out/r2g_c5_n2.json 2> gpurun_out/r2g_c5_n2.err; echo "c5 rc=$?"
timeout 600 python bench.py --workload c5 --c5-videos 32 --steps 2 > gpurun_out/r2g_c5_n1.json 2> gpurun_out/r2g_c5_n1.err; echo "c5-1 rc=$?"
