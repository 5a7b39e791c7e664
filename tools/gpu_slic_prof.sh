#!/bin/bash
set -u
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/slic_launches.csv python tools/slic_run.py > gpurun_out/slic_ncu.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/slic_launches.csv | head -10
