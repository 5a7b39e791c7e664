#!/bin/bash
# round 2: five CTAs of k_q8_front per SM (48 registers, no spills) against four
set -u
out=gpurun_out
mkdir -p $out
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline'].get('kernels') or {}
print('$2', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), {n: round(v['ms_per_step'],4) for n, v in k.items() if isinstance(v, dict)})"; }
V=$PWD/depth_completion_mt_b200/variants/libdcmt_front5.so
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2w_4.json 2>> $out/r2w.err; show $out/r2w_4.json "4 CTAs:"
DCMT_LIB=$V timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2w_5.json 2>> $out/r2w.err; show $out/r2w_5.json "5 CTAs auto:"
DCMT_LIB=$V DCMT_FRONT_TILE_W=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2w_5w0.json 2>> $out/r2w.err; show $out/r2w_5w0.json "5 CTAs, 152 wide:"
DCMT_LIB=$V DCMT_FRONT_TILE_W=48 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2w_5w48.json 2>> $out/r2w.err; show $out/r2w_5w48.json "5 CTAs, 48 wide:"
DCMT_LIB=$V timeout 300 python bench.py --input u16 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2w_5u16.json 2>> $out/r2w.err; show $out/r2w_5u16.json "5 CTAs auto u16:"
DCMT_LIB=$V timeout 300 python bench.py --rows 1024 --cols 2048 --frames 192 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2w_5big.json 2>> $out/r2w.err; show $out/r2w_5big.json "5 CTAs auto 1024x2048:"
