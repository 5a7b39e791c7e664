#!/bin/bash
set -u
for mpx in 2 4 8 16 32; do
  DCMT_HOST_CHUNK_MPX=$mpx timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_$mpx.json 2>> gpurun_out/e2e.err
  python - <<PY
import json
d=json.load(open("gpurun_out/e2e_$mpx.json"))
print("host chunk $mpx Mpx: e2e f32", round(d["e2e"]["value"]), "e2e u16", round(d["e2e_u16_input"]["value"]))
PY
done
