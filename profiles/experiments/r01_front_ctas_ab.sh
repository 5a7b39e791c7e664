#!/bin/bash
# A/B: k_q8_front with 3 or 4 CTAs per SM on lower tiles (DCMT_FRONT_TILE_H) against the 2-CTA / 88-row default
set -u
out=gpurun_out; mkdir -p $out; : > $out/front3.txt
run() {  # label lib h
  local label=$1 lib=$2 h=$3
  ( if [ -n "$lib" ]; then export DCMT_LIB=$PWD/depth_completion_mt_b200/ab/libdcmt_$lib.so; fi
    if [ "$h" != 0 ]; then export DCMT_FRONT_TILE_H=$h; fi
    timeout 300 python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu-baseline 2>> $out/front3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label', 'frames/s', round(d['value']), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items() if isinstance(v,dict)}, 'golden', d['validation']['golden_sha256_match'])" ) | tee -a $out/front3.txt
}
run "old default 2x512 h88" f2x512 0
run "new default 4x256 (chooser)" "" 0
run "4x224 h44" f4x224 44
run "4x288 h44" f4x288 44
run "4x320 h44" f4x320 44
run "5x192 h32" f5x192 32
run "5x224 h32" f5x224 32
run "new default again" "" 0
run "old default again" f2x512 0
