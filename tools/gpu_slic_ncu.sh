#!/bin/bash
set -u
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_slic_assign -s 12 -c 1 -o gpurun_out/slic_assign -f python tools/slic_run.py > gpurun_out/slic_ncu_full.log 2>&1; echo "rc=$?"
