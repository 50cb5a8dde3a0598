#!/bin/bash
# 8-GPU evidence: bench.py under the driver's launch line at N = 8 and N = 4 (dp = parent training, batch 12 per GPU, one
# NCCL allreduce per step; headline = inference replicas)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c10_gpus.txt
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --skip gpu_reference ) > gpurun_out/c10_bench_8gpu.json 2>gpurun_out/c10_bench_8gpu.err
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 5 --skip gpu_reference,parity ) > gpurun_out/c10_bench_4gpu.json 2>gpurun_out/c10_bench_4gpu.err
( NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 5 --warmup 3 --workload parent480 ) > gpurun_out/c10_bench_parent_8gpu.json 2>gpurun_out/c10_nccl_info.txt
grep -i "nvls\|Channel\|algo" gpurun_out/c10_nccl_info.txt | head -8
tail -c 2500 gpurun_out/c10_bench_8gpu.json; echo; tail -c 300 gpurun_out/c10_bench_8gpu.err; tail -c 1200 gpurun_out/c10_bench_4gpu.json
