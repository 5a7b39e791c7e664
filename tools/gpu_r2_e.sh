#!/bin/bash
# round 2: the float32 dictionary path (DCMT_PATH_RANK) -- parity on the GPU and throughput next to the generic pipeline
set -u
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests/test_fused_q8.py tests/test_guided_stereo_parity.py -x -q -m gpu > $out/r2e_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $out/r2e_pytest.log
for cfg in "lidar_only rank" "lidar_only generic" "guided rank" "guided generic"; do
  set -- $cfg
  timeout 600 python bench.py --workload $1 --input float --path $2 --frames 512 --steps 10 --warmup 3 --no-cpu-baseline > $out/r2e_float_$1_$2.json 2>> $out/r2e.err
  python - <<PY
import json
d=json.load(open("$out/r2e_float_$1_$2.json"))
print("$1 float input, path $2: frames/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ceiling", d["e2e"].get("copy_ceiling"))
PY
done
tail -3 $out/r2e.err
