#!/bin/bash
# round 2, closing one-GPU session after the front's tile width and the band SLIC kernel: what profiles/README.md quotes for the
# lidar-only path, the float path, the sweep and the rows either side of the path (guided / stereo kernels did not change:
# profiles/r02_final_bench_guided.json, _stereo.json stand)
set -u
out=gpurun_out
p=$out/final2
mkdir -p $p
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $p/gpu.txt 2>&1
timeout 1800 python -m pytest tests -q -m gpu > $p/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $p/pytest_gpu.log
tail -3 $p/pytest_gpu.log
timeout 900 python bench.py > $p/bench.json 2> $p/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $p/bench_reference.json 2>> $p/bench.err; echo "ref rc=$?"
timeout 600 python bench.py --input u16 --no-cpu-baseline > $p/bench_u16_input.json 2>> $p/bench.err; echo "u16 rc=$?"
timeout 600 python bench.py --workload lidar_only --input float --path rank --frames 512 --steps 10 --warmup 3 --no-cpu-baseline > $p/bench_float_rank.json 2>> $p/bench.err
timeout 900 python bench.py --workload stereo_chain --frames 256 --steps 5 --warmup 3 --no-cpu-baseline > $p/bench_stereo_chain.json 2>> $p/bench.err; echo "chain rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$p/bench*.json")):
    try:
        d=json.load(open(f)); e=d.get("e2e") or {}
        print(f.split("/")[-1], "value", round(d["value"]), "frac", d.get("roofline",{}).get("frac") and round(d["roofline"]["frac"],4), "e2e", e.get("value") and round(e["value"]), "ceil", e.get("copy_ceiling") and round(e["copy_ceiling"]), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as exc: print(f, "FAILED", exc)
PY
: > $p/sweep.jsonl
for shape in "352 1216 1024" "512 1760 512" "1024 2048 192" "2048 4096 48"; do
  set -- $shape
  for dens in 0.01 0.02 0.05 0.1 0.2; do
    timeout 300 python bench.py --rows $1 --cols $2 --frames $3 --density $dens --steps 10 --warmup 3 --no-e2e --no-cpu-baseline >> $p/sweep.jsonl 2>> $p/sweep.err || echo "{\"failed\": \"$1x$2 $dens\"}" >> $p/sweep.jsonl
  done
done
python - > $p/sweep.txt <<'PY'
import json
for l in open("gpurun_out/final2/sweep.jsonl"):
    d = json.loads(l)
    if "failed" in d: print(d); continue
    c = d["config"]
    print(c["rows"], c["cols"], c["valid_density"], "frames/s", round(d["value"]), "Mpx/s", round(d["value"]*c["rows"]*c["cols"]/1e6), "frac", round(d["roofline"]["frac"],4), d["validation"]["replicas_equal"])
PY
cat $p/sweep.txt
python tools/bench_rows.py --reps 10 > $p/rows_bench.jsonl 2>> $p/bench.err; cut -c1-200 $p/rows_bench.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $p/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $p/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_q8_(front|tail)' -s 6 -c 2 -o $p/q8 -f \
  python bench.py --steps 1 --warmup 3 --frames 512 --no-e2e --no-cpu-baseline > $p/ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/q8_phase_profile.py 158 > $p/phase_cycles.txt 2>&1
python tools/q8_phase_profile.py 1 >> $p/phase_cycles.txt 2>&1
timeout 200 python tools/fuzz_gpu.py 120 23 > $p/fuzz_gpu.txt 2>&1; tail -2 $p/fuzz_gpu.txt
ls $p
