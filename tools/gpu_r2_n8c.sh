#!/bin/bash
# round 2, 8-GPU box, closing kernels (the front's own tile width): BASELINE configs[1] at N = 1 and N = 8 on one box
set -u
out=gpurun_out
mkdir -p $out
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/r2c_lidar_1gpu.json 2>> $out/r2c.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > $out/r2c_lidar_8gpu.json 2>> $out/r2c.err
python - <<PY
import json
for n in (1, 8):
    try:
        d=json.load(open("$out/r2c_lidar_%dgpu.json" % n)); e=d.get("e2e") or {}
        print("lidar N=%d value" % n, round(d["value"]), "e2e", e.get("value") and round(e["value"]), "ceiling", e.get("copy_ceiling") and round(e["copy_ceiling"]), "ranks", d["validation"]["ranks"])
    except Exception as exc:
        print("N=%d FAILED" % n, exc)
PY
tail -2 $out/r2c.err
