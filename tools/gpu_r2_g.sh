#!/bin/bash
# round 2: dictionary path after the encode / compact / decode rework; chain parity + chain bench; compute-sanitizer on a small case
set -u
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests/test_project.py tests/test_fused_q8.py tests/test_stereo_chain.py tests/test_guided_stereo_parity.py -x -q -m gpu > $out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $out/r2g_pytest.log
for cfg in "lidar_only rank" "guided rank"; do
  set -- $cfg
  timeout 600 python bench.py --workload $1 --input float --path $2 --frames 512 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/r2g_float_$1_$2.json 2>> $out/r2g.err
  python - <<PY
import json
d=json.load(open("$out/r2g_float_$1_$2.json"))
print("$1 float input, path $2: frames/s", round(d["value"]))
PY
done
timeout 900 python bench.py --workload stereo_chain --frames 256 --steps 5 --warmup 3 --no-cpu-baseline > $out/r2g_chain.json 2>> $out/r2g.err; echo "chain rc=$?"
python - <<PY
import json
d=json.load(open("$out/r2g_chain.json"))
print("stereo_chain frames/s", round(d["value"]), "frac", round(d["roofline"]["frac"],4))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/r2g_launches.csv python bench.py --workload lidar_only --input float --path rank --frames 512 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/r2g_ncu.log 2>&1
python tools/summarize_launches.py $out/r2g_launches.csv 2>/dev/null | head -8
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $out/r2g_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $out/r2g_memcheck.log
tail -3 $out/r2g.err
