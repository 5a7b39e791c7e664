#!/bin/bash
# round 2, first GPU session: parity of the shared-work median, A/B against the old median and other CTA sizes, phase profile
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_fused_q8.py tests/test_completion_parity.py -x -q -m gpu > $out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $out/r2a_pytest.log
for v in "" oldmed t384 t448 ""; do
  if [ -z "$v" ]; then unset DCMT_LIB; else export DCMT_LIB=$PWD/depth_completion_mt_b200/variants/libdcmt_$v.so; fi
  timeout 300 python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2a_var_$v.json 2>> $out/r2a_var.err
  python - <<PY
import json
d=json.load(open("$out/r2a_var_$v.json"))
print("variant '$v' frames/s", round(d["value"]), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)}, d["validation"]["golden_sha256_match"])
PY
done
unset DCMT_LIB
python tools/q8_phase_profile.py 158 > $out/r2a_phase.txt 2>&1
python tools/q8_phase_profile.py 1 >> $out/r2a_phase.txt 2>&1
cat $out/r2a_phase.txt
