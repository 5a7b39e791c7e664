#!/bin/bash
# BASELINE configs[4]: resolution / density sweep of the lidar-only hot path on one GPU (device-resident, no CPU leg)
set -u
out=gpurun_out/sweep.jsonl
: > $out
for shape in "352 1216 1024" "512 1760 512" "1024 2048 192" "2048 4096 48"; do
  set -- $shape
  for dens in 0.01 0.02 0.05 0.1 0.2; do
    timeout 300 python bench.py --rows $1 --cols $2 --frames $3 --density $dens --steps 10 --warmup 3 --no-e2e --no-cpu-baseline >> $out 2>> gpurun_out/sweep.err || echo "{\"failed\": \"$1x$2 $dens\"}" >> $out
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/sweep.jsonl"):
    d = json.loads(l)
    if "failed" in d: print(d); continue
    c = d["config"]
    print(c["rows"], c["cols"], c["valid_density"], "frames/s", round(d["value"]), "Mpx/s", round(d["value"]*c["rows"]*c["cols"]/1e6), "frac", round(d["roofline"]["frac"],4), d["validation"]["replicas_equal"])
PY
