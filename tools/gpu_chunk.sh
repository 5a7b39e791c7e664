#!/bin/bash
set -u
out=gpurun_out
for c in 79 158 316 512 1024; do
  DCMT_FUSED_CHUNK=$c timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 > $out/chunk_$c.json 2> $out/chunk_$c.err
  python - <<PY
import json
d=json.load(open("$out/chunk_$c.json"))
print("chunk $c value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)})
PY
done
