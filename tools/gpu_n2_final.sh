#!/bin/bash
# 2 x B200 under torchrun, launched the way the driver does: our arm and the reference CPU arm
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r01_final_bench_2gpu.json 2> gpurun_out/r01_final_bench_2gpu.err; echo "n2 rc=$?"
cut -c1-200 gpurun_out/r01_final_bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r01_final_bench_reference.json 2>> gpurun_out/r01_final_bench_2gpu.err; echo "ref n2 rc=$?"
cut -c1-200 gpurun_out/r01_final_bench_reference.json
