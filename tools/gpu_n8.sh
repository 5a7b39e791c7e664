#!/bin/bash
set -u
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r01_bench_n$N.json 2> gpurun_out/r01_bench_n$N.err; echo "n$N rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r01_bench_n$N.json"))
print("N=$N value", round(d["value"]), "per-gpu", round(d["value"]/$N), "e2e", round(d["e2e"]["value"]), "e2e_u16", round(d["e2e_u16_input"]["value"]), d["validation"]["checksums_equal_across_ranks"], d["validation"]["golden_sha256_match"], [round(x,1) for x in d["validation"]["ms_per_rank"]])
PY
tail -3 gpurun_out/r01_bench_n$N.err
