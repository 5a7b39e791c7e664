#!/bin/bash
# A/B of the overlapped front / tail mode of the fused path (DCMT_OVERLAP, DCMT_OVERLAP_CHUNK, DCMT_OVERLAP_SOLO)
set -u
out=gpurun_out; mkdir -p $out; : > $out/overlap.txt
timeout 900 python -m pytest tests/test_fused_q8.py tests/test_completion_parity.py -x -q -m gpu 2>&1 | tail -2 | tee -a $out/overlap.txt
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu-baseline ${EXTRA:-} 2>> $out/overlap.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label', 'frames/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'golden', d['validation']['golden_sha256_match'])" | tee -a $out/overlap.txt
}
run "overlap=0" DCMT_OVERLAP=0
run "overlap=1 default-chunk" DCMT_OVERLAP=1
run "overlap=1 chunk=64" DCMT_OVERLAP=1 DCMT_OVERLAP_CHUNK=64
run "overlap=1 chunk=103" DCMT_OVERLAP=1 DCMT_OVERLAP_CHUNK=103
run "overlap=1 chunk=256" DCMT_OVERLAP=1 DCMT_OVERLAP_CHUNK=256
run "overlap=1 chunk=342" DCMT_OVERLAP=1 DCMT_OVERLAP_CHUNK=342
run "overlap=1 default-chunk solo=0" DCMT_OVERLAP=1 DCMT_OVERLAP_SOLO=0
run "overlap=1 chunk=64 solo=0" DCMT_OVERLAP=1 DCMT_OVERLAP_SOLO=0 DCMT_OVERLAP_CHUNK=64
EXTRA="--input u16" run "u16 overlap=0" DCMT_OVERLAP=0
EXTRA="--input u16" run "u16 overlap=1" DCMT_OVERLAP=1
