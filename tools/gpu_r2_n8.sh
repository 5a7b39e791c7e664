#!/bin/bash
# round 2, one 8-GPU box: BASELINE configs[1] at N = 2, 4, 8 (e2e next to its copy-only ceiling), configs[2] / [3] and the
# stereo chain at N = 8, the sweep corners of configs[4] at N = 8, and the single-process multi-device host entry point
set -u
out=gpurun_out
mkdir -p $out
run() {  # N tag args...
  local n=$1 tag=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $n "$@" > $out/r2n_${tag}_${n}gpu.json 2>> $out/r2n.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r2n_${tag}_${n}gpu.json"))
    e=d.get("e2e") or {}
    print("$tag N=$n value", round(d["value"]), "e2e", e.get("value") and round(e["value"]), "ceiling", e.get("copy_ceiling") and round(e["copy_ceiling"]),
          "f32 e2e", (d.get("e2e_f32_input") or {}).get("value"), "ranks", d["validation"]["ranks"])
except Exception as exc:
    print("$tag N=$n FAILED", exc)
PY
}
nvidia-smi -L | head -8
nvidia-smi topo -m > $out/r2n_topo.txt 2>&1
for n in 2 4 8; do run $n lidar --steps 20 --warmup 5 --no-cpu-baseline; done
run 8 guided --workload guided --frames 256 --steps 10 --warmup 3 --no-cpu-baseline
run 8 stereo --workload stereo --frames 512 --steps 10 --warmup 3 --no-cpu-baseline
run 8 chain --workload stereo_chain --frames 256 --steps 5 --warmup 3 --no-cpu-baseline
run 8 float --workload lidar_only --input float --path rank --frames 512 --steps 10 --warmup 3 --no-cpu-baseline
run 8 sweep_352x1216_p01 --density 0.01 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
run 8 sweep_2048x4096_p01 --rows 2048 --cols 4096 --density 0.01 --frames 64 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
run 8 sweep_2048x4096_p20 --rows 2048 --cols 4096 --density 0.2 --frames 64 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
# one process, one thread, all 8 GPUs through dcmt_img_completion_u16_host_multi
timeout 900 python bench.py --host-multi --steps 10 --warmup 3 --no-cpu-baseline > $out/r2n_host_multi.json 2>> $out/r2n.err
python - <<PY
import json
d=json.load(open("$out/r2n_host_multi.json"))
print("single process: e2e one GPU", round(d["e2e"]["value"]), "host_multi over", d["e2e_host_multi"]["devices"], "GPUs:", round(d["e2e_host_multi"]["value"]))
PY
timeout 600 python -m pytest tests/test_abi.py -x -q -m gpu -k multi > $out/r2n_pytest_multi.log 2>&1; tail -2 $out/r2n_pytest_multi.log
tail -5 $out/r2n.err
