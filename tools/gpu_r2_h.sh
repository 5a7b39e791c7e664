#!/bin/bash
# round 2: float path after the sort / encode rework, launch lists of the float path and the stereo chain,
# ncu --set full of the stereo and SLIC kernels (summaries asked for by the round-1 review)
set -u
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests/test_abi.py tests/test_fused_q8.py tests/test_stereo_chain.py -x -q -m gpu > $out/r2h_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $out/r2h_pytest.log
for cfg in "lidar_only rank" "guided rank"; do
  set -- $cfg
  timeout 600 python bench.py --workload $1 --input float --path $2 --frames 512 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/r2h_float_$1_$2.json 2>> $out/r2h.err
  python - <<PY
import json
d=json.load(open("$out/r2h_float_$1_$2.json"))
print("$1 float input, path $2: frames/s", round(d["value"]))
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/r2h_launches_float.csv python bench.py --workload lidar_only --input float --path rank --frames 512 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/r2h_ncu.log 2>&1
python tools/summarize_launches.py $out/r2h_launches_float.csv 2>/dev/null | head -6
timeout 900 python bench.py --workload stereo_chain --frames 256 --steps 5 --warmup 3 --no-cpu-baseline > $out/r2h_chain.json 2>> $out/r2h.err; echo "chain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r2h_launches_chain.csv python bench.py --workload stereo_chain --frames 128 --steps 1 --warmup 3 --no-cpu-baseline > $out/r2h_ncu2.log 2>&1
python tools/summarize_launches.py $out/r2h_launches_chain.csv 2>/dev/null | head -14
python tools/slic_run.py 1; python tools/slic_run.py 64; python tools/slic_run.py 256
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_slic_(assign|prepare)' -s 8 -c 2 -o $out/r2h_slic -f python tools/slic_run.py 64 > $out/r2h_slic_ncu.log 2>&1; echo "slic ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stereo_refine -s 3 -c 1 -o $out/r2h_stereo -f python bench.py --workload stereo --frames 128 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2h_stereo_ncu.log 2>&1; echo "stereo ncu rc=$?"
tail -3 $out/r2h.err
