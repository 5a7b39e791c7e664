#!/bin/bash
# round 2: dictionary path throughput after moving the compaction off the sort CTA; batched projection parity
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_project.py tests/test_fused_q8.py -x -q -m gpu -k "project or rank" > $out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $out/r2f_pytest.log
for cfg in "lidar_only rank" "guided rank"; do
  set -- $cfg
  timeout 600 python bench.py --workload $1 --input float --path $2 --frames 512 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/r2f_float_$1_$2.json 2>> $out/r2f.err
  python - <<PY
import json
d=json.load(open("$out/r2f_float_$1_$2.json"))
print("$1 float input, path $2: frames/s", round(d["value"]))
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/r2f_launches.csv python bench.py --workload lidar_only --input float --path rank --frames 512 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/r2f_ncu.log 2>&1
python tools/summarize_launches.py $out/r2f_launches.csv 2>/dev/null | tail -15
tail -3 $out/r2f.err
