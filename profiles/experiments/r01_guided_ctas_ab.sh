#!/bin/bash
# A/B: per-label guided front with two CTAs per SM on lower tiles (DCMT_GTL, DCMT_GUIDED_CTAS, DCMT_GUIDED_TILE_H)
set -u
out=gpurun_out; mkdir -p $out; : > $out/guided_ab.txt
timeout 600 python -m pytest tests/test_guided_stereo_parity.py -x -q -m gpu 2>&1 | tail -1 | tee -a $out/guided_ab.txt
run() {  # label lib h
  local label=$1 lib=$2 h=$3
  ( if [ -n "$lib" ]; then export DCMT_LIB=$PWD/depth_completion_mt_b200/ab/libdcmt_$lib.so; fi
    if [ "$h" != 0 ]; then export DCMT_GUIDED_TILE_H=$h; fi
    timeout 300 python bench.py --workload guided --frames 256 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>> $out/guided_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label', 'frames/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'replicas_equal', d['validation']['replicas_equal'], 'checksum', d['validation'].get('checksum', d['validation'].get('checksums_equal_across_ranks')))" ) | tee -a $out/guided_ab.txt
}
run "default 1x1024 h88" "" 0
run "2x512 h44" g2x512 44
run "2x512 h59" g2x512 59
run "2x512 chooser" g2x512 0
run "1x512 h88" g1x512 0
run "2x384 h44" g2x384 44
run "1x1024 h44" "" 44
run "default again" "" 0
