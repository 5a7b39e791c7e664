#!/bin/bash
# round 2: tile width of the guided front
set -u
out=gpurun_out
mkdir -p $out
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline'].get('kernels') or {}
print('$2', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), {n: round(v['ms_per_step'],4) for n, v in k.items() if isinstance(v, dict)})"; }
for w in 0 112 176 48; do
  DCMT_GUIDED_TILE_W=$w timeout 300 python bench.py --workload guided --frames 256 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2v_guided_$w.json 2>> $out/r2v.err; show $out/r2v_guided_$w.json "guided width $w:"
  DCMT_GUIDED_TILE_W=$w timeout 300 python bench.py --workload guided --input float --path rank --frames 256 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2v_guidedf_$w.json 2>> $out/r2v.err; show $out/r2v_guidedf_$w.json "guided float width $w:"
done
