#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_guided_stereo_parity.py tests/test_slic.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --workload guided --frames 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/guided_bench.json 2>> gpurun_out/guided.err
python -c "
import json
d=json.load(open('gpurun_out/guided_bench.json')); print('guided frames/s', round(d['value']), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']))"
timeout 300 python tools/bench_rows.py --reps 10 2>> gpurun_out/guided.err | tail -1 | cut -c1-200
