#!/bin/bash
# round 2: full GPU suite, median run-length A/B, float path and chain after the encode / projection fixes
set -u
out=gpurun_out
mkdir -p $out
timeout 1800 python -m pytest tests -x -q -m gpu > $out/r2i_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $out/r2i_pytest.log
for ml in 0 12 24 32 46; do
  DCMT_MED_LEN=$ml timeout 300 python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2i_med_$ml.json 2>> $out/r2i.err
  python - <<PY
import json
d=json.load(open("$out/r2i_med_$ml.json"))
print("DCMT_MED_LEN=$ml frames/s", round(d["value"]), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)}, d["validation"]["golden_sha256_match"])
PY
done
for cfg in "lidar_only rank" "guided rank"; do
  set -- $cfg
  timeout 600 python bench.py --workload $1 --input float --path $2 --frames 512 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/r2i_float_$1_$2.json 2>> $out/r2i.err
  python - <<PY
import json
d=json.load(open("$out/r2i_float_$1_$2.json"))
print("$1 float input, path $2: frames/s", round(d["value"]))
PY
done
timeout 900 python bench.py --workload stereo_chain --frames 256 --steps 5 --warmup 3 --no-cpu-baseline > $out/r2i_chain.json 2>> $out/r2i.err; echo "chain rc=$?"
python - <<PY
import json
d=json.load(open("$out/r2i_chain.json"))
print("stereo_chain frames/s", round(d["value"]))
PY
tail -3 $out/r2i.err
