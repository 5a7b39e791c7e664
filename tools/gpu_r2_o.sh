#!/bin/bash
# round 2: where the batched projection and the band SLIC kernel spend their time
set -u
out=gpurun_out
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/r2o_launches_proj.csv python tools/proj_run.py 128 > $out/r2o_ncu1.log 2>&1
python tools/summarize_launches.py $out/r2o_launches_proj.csv 2>/dev/null | head -8
DCMT_SLIC_BAND_MIN_FRAMES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_slic_assign_band' -s 4 -c 1 -o $out/r2o_slic_band -f python tools/slic_run.py 64 > $out/r2o_slic_ncu.log 2>&1; echo "slic band ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_project_' -s 8 -c 4 -o $out/r2o_proj -f python tools/proj_run.py 128 > $out/r2o_proj_ncu.log 2>&1; echo "proj ncu rc=$?"
