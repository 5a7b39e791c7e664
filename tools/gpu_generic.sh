#!/bin/bash
set -u
for c in 20 80 320; do
  DCMT_GENERIC_CHUNK=$c timeout 300 python bench.py --path generic --frames 640 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/gen.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('generic chunk $c frames/s', round(d['value']))"
done
timeout 300 python bench.py --workload guided --frames 256 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/gen.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('guided frames/s', round(d['value']))"
timeout 300 python bench.py --workload stereo --frames 512 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/gen.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('stereo frames/s', round(d['value']))"
