#!/bin/bash
# round 2: median run length sweep on the current tail
set -u
out=gpurun_out
mkdir -p $out
for ml in 0 20 28 32 36 40 0; do
  DCMT_MED_LEN=$ml timeout 300 python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2l_med_$ml.json 2>> $out/r2l.err
  python - <<PY
import json
d=json.load(open("$out/r2l_med_$ml.json"))
print("DCMT_MED_LEN=$ml frames/s", round(d["value"]), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)}, d["validation"]["golden_sha256_match"])
PY
done
python tools/slic_run.py 1; python tools/slic_run.py 64; python tools/slic_run.py 256
tail -3 $out/r2l.err
