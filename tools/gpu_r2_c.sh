#!/bin/bash
# round 2: full GPU suite after the API rework, then A/B of the compare-exchange pipe split
set -u
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -x -q -m gpu > $out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $out/r2c_pytest.log
for v in "" o13 o12 o23 o11 ""; do
  if [ -z "$v" ]; then unset DCMT_LIB; else export DCMT_LIB=$PWD/depth_completion_mt_b200/variants/libdcmt_$v.so; fi
  timeout 300 python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2c_var_$v.json 2>> $out/r2c_var.err
  python - <<PY
import json
d=json.load(open("$out/r2c_var_$v.json"))
print("variant '$v' frames/s", round(d["value"]), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)}, d["validation"]["golden_sha256_match"])
PY
done
unset DCMT_LIB
python tools/q8_phase_profile.py 158 > $out/r2c_phase.txt 2>&1
tail -11 $out/r2c_phase.txt
