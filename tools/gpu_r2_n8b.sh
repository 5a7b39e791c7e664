#!/bin/bash
# round 2, 8-GPU box, final kernels: BASELINE configs[1] at N = 1, 2, 4, 8 on ONE box (e2e next to its copy-only ceiling), the
# stereo chain at N = 8 after the projection fix, the multi-device host entry point, the multi-GPU ABI test
set -u
out=gpurun_out
mkdir -p $out
run() {  # N tag args...
  local n=$1 tag=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 900 python bench.py "$@" > $out/r2p_${tag}_${n}gpu.json 2>> $out/r2p.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n "$@" > $out/r2p_${tag}_${n}gpu.json 2>> $out/r2p.err
  fi
  python - <<PY
import json
try:
    d=json.load(open("$out/r2p_${tag}_${n}gpu.json"))
    e=d.get("e2e") or {}
    print("$tag N=$n value", round(d["value"]), "e2e", e.get("value") and round(e["value"]), "ceiling", e.get("copy_ceiling") and round(e["copy_ceiling"]),
          "f32 e2e", (d.get("e2e_f32_input") or {}).get("value"), "ranks", d["validation"]["ranks"])
except Exception as exc:
    print("$tag N=$n FAILED", exc)
PY
}
for n in 1 2 4 8; do run $n lidar --steps 20 --warmup 5 --no-cpu-baseline; done
run 1 chain --workload stereo_chain --frames 256 --steps 5 --warmup 3 --no-cpu-baseline
run 8 chain --workload stereo_chain --frames 256 --steps 5 --warmup 3 --no-cpu-baseline
timeout 900 python bench.py --host-multi --steps 10 --warmup 3 --no-cpu-baseline > $out/r2p_host_multi.json 2>> $out/r2p.err
python - <<PY
import json
d=json.load(open("$out/r2p_host_multi.json"))
print("single process: e2e one GPU", round(d["e2e"]["value"]), "host_multi over", d["e2e_host_multi"]["devices"], "GPUs:", round(d["e2e_host_multi"]["value"]))
PY
timeout 600 python -m pytest tests/test_abi.py tests/test_project.py -x -q -m gpu > $out/r2p_pytest.log 2>&1; tail -2 $out/r2p_pytest.log
python tools/bench_rows.py --reps 10 2>/dev/null | cut -c1-170 | sed -n 2,3p
tail -3 $out/r2p.err
