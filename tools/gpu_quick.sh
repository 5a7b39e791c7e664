#!/bin/bash
# quick GPU check: parity tests (bounded), default bench without the CPU leg, optional A/B env toggles
set -u
tag=${1:-q}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 $out/${tag}_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$out/${tag}_bench.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)}, d["validation"])
PY
DCMT_NO_TMA=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e > $out/${tag}_bench_notma.json 2>> $out/${tag}_bench.err; echo "bench notma rc=$?"
python - <<PY
import json
d=json.load(open("$out/${tag}_bench_notma.json"))
print("NO_TMA value", d["value"], "ms/step", d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)})
PY
tail -5 $out/${tag}_bench.err
