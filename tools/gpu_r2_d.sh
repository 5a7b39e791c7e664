#!/bin/bash
# round 2: the new bench line (copy ceilings, latency, u16 e2e headline) on one GPU + the GPU suite incl. the cv::Mat shim
set -u
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -x -q -m gpu > $out/r2d_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $out/r2d_pytest.log
timeout 900 python bench.py > $out/r2d_bench.json 2> $out/r2d_bench.err; echo "bench rc=$?"
tail -3 $out/r2d_bench.err
python - <<PY
import json
d=json.load(open("$out/r2d_bench.json"))
print("value", round(d["value"]), "frac", round(d["roofline"]["frac"],4), "issue", d["roofline"]["issue"] and round(d["roofline"]["issue"]["frac"],3))
print("e2e", d["e2e"]["value"], "ceiling", d["e2e"]["copy_ceiling"], "f32", d["e2e_f32_input"]["value"], d["e2e_f32_input"]["copy_ceiling"])
print("latency", d.get("latency"))
print("cpu", d.get("cpu_baseline"))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/r2d_bench_reference.json 2>> $out/r2d_bench.err; echo "ref rc=$?"
