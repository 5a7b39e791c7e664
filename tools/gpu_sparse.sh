#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_fused_q8.py -x -q -m gpu 2>&1 | tail -3
for dens in 0.005 0.01 0.02 0.05; do
  timeout 300 python bench.py --density $dens --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/sparse_$dens.json 2>> gpurun_out/sparse.err
  python - <<PY
import json
d=json.load(open("gpurun_out/sparse_$dens.json"))
print("density $dens frames/s", round(d["value"]), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)}, d["validation"]["replicas_equal"])
PY
done
timeout 300 python bench.py --rows 2048 --cols 4096 --frames 48 --density 0.01 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/sparse_big.json 2>> gpurun_out/sparse.err
python -c "
import json
d=json.load(open('gpurun_out/sparse_big.json')); print('2048x4096 1% frames/s', round(d['value'],1))"
