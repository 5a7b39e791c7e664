#!/bin/bash
set -u
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stereo_refine -s 3 -c 1 -o gpurun_out/stereo -f python bench.py --workload stereo --frames 128 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/stereo_ncu.log 2>&1; echo "rc=$?"
