#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_guided_stereo_parity.py tests/test_slic.py -x -q -m gpu 2>&1 | tail -3
for lab in grid slic; do for mode in 1 0; do
  DCMT_GUIDED_PER_LABEL=$mode timeout 300 python bench.py --workload guided --labels $lab --frames 256 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/guided.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('labels $lab per_label $mode frames/s', round(d['value']))"
done; done
