#!/bin/bash
# round 2: word-level hole list in the tail, band-based SLIC assignment
set -u
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests/test_fused_q8.py tests/test_slic.py tests/test_completion_parity.py tests/test_reference_build.py tests/test_guided_stereo_parity.py -x -q -m gpu > $out/r2k_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $out/r2k_pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2k_bench_$i.json 2>> $out/r2k.err
python - <<PY
import json
d=json.load(open("$out/r2k_bench_$i.json"))
print("frames/s", round(d["value"]), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)}, d["validation"]["golden_sha256_match"])
PY
done
python tools/slic_run.py 1; python tools/slic_run.py 8; python tools/slic_run.py 64; python tools/slic_run.py 256
python tools/q8_phase_profile.py 158 2>&1 | tail -10
python tools/bench_rows.py --reps 10 > $out/r2k_rows.jsonl 2>> $out/r2k.err; cut -c1-230 $out/r2k_rows.jsonl
tail -3 $out/r2k.err
