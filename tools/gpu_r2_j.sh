#!/bin/bash
# round 2: tail with 8-byte vertical maxima, SLIC with the integer window test
set -u
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests/test_fused_q8.py tests/test_slic.py tests/test_completion_parity.py tests/test_reference_build.py -x -q -m gpu > $out/r2j_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $out/r2j_pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2j_bench_$i.json 2>> $out/r2j.err
python - <<PY
import json
d=json.load(open("$out/r2j_bench_$i.json"))
print("frames/s", round(d["value"]), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)}, d["validation"]["golden_sha256_match"])
PY
done
python tools/slic_run.py 1; python tools/slic_run.py 64
python tools/q8_phase_profile.py 158 2>&1 | tail -10
tail -3 $out/r2j.err
