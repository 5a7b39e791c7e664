#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_fused_q8.py -x -q -m gpu 2>&1 | tail -2
for m in 1 0; do
  DCMT_FRONT_STREAM=$m timeout 300 python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/stream.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('front_stream $m frames/s', round(d['value']), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items() if isinstance(v,dict)}, d['validation']['golden_sha256_match'])"
done
DCMT_FRONT_STREAM=1 timeout 300 python bench.py --input u16 --steps 15 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/stream.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('front_stream 1 u16 frames/s', round(d['value']), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items() if isinstance(v,dict)})"
