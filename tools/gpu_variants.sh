#!/bin/bash
set -u
for v in "" t448 t576 t640 f384 f640 f256; do
  if [ -z "$v" ]; then unset DCMT_LIB; else export DCMT_LIB=$PWD/depth_completion_mt_b200/variants/libdcmt_$v.so; fi
  timeout 300 python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/var_$v.json 2>> gpurun_out/var.err
  python - <<PY
import json
d=json.load(open("gpurun_out/var_$v.json"))
print("variant '$v' frames/s", round(d["value"]), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)}, d["validation"]["golden_sha256_match"])
PY
done
