#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_fused_q8.py -x -q -m gpu 2>&1 | tail -2
for shape in "1024 2048 192" "2048 4096 48" "512 1760 512"; do
  set -- $shape
  timeout 300 python bench.py --rows $1 --cols $2 --frames $3 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/n2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']; print(c['rows'],c['cols'],'frames/s',round(d['value']),'Mpx/s',round(d['value']*c['rows']*c['cols']/1e6))"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r01k_bench_n2.json 2> gpurun_out/r01k_bench_n2.err; echo "n2 rc=$?"
cat gpurun_out/r01k_bench_n2.json | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r01k_bench_ref_n2.json 2>> gpurun_out/r01k_bench_n2.err; echo "ref n2 rc=$?"
cut -c1-300 gpurun_out/r01k_bench_ref_n2.json
