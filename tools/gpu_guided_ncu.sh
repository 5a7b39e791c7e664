#!/bin/bash
set -u
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_q8_guided -s 3 -c 1 -o gpurun_out/guided_q8 -f python bench.py --workload guided --frames 79 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/guided_ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out/guided_q8.ncu-rep
