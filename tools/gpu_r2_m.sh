#!/bin/bash
# round 2: default bench line (copy ceiling before / after each e2e leg), front CTA variants 3 x 320 / 3 x 256
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python bench.py > $out/r2m_bench.json 2> $out/r2m_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$out/r2m_bench.json"))
print("value", round(d["value"]), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), "ceil", round(d["e2e"]["copy_ceiling"]), "f32", round(d["e2e_f32_input"]["value"]), round(d["e2e_f32_input"]["copy_ceiling"]), "cpu", d["cpu_baseline"]["value"], "lat", d["latency"]["u16_host"], d["latency"]["reference_single_thread_ms"])
PY
for v in f3x320 f3x256 ""; do
  if [ -z "$v" ]; then unset DCMT_LIB; else export DCMT_LIB=$PWD/depth_completion_mt_b200/variants/libdcmt_$v.so; fi
  timeout 300 python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2m_var_$v.json 2>> $out/r2m.err
  python - <<PY
import json
d=json.load(open("$out/r2m_var_$v.json"))
print("variant '$v' frames/s", round(d["value"]), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)}, d["validation"]["golden_sha256_match"])
PY
done
unset DCMT_LIB
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
