#!/bin/bash
# one ncu --set full capture of the fused kernels at the bench's own chunk size (512 frames per launch), after a plain run
set -u
tag=${1:-r2}
frames=${2:-512}
out=gpurun_out
mkdir -p $out
timeout 600 python bench.py --steps 1 --warmup 3 --frames $frames --no-e2e --no-cpu-baseline > $out/${tag}_plain.json 2> $out/${tag}_plain.err; echo "plain rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_q8_(front|tail)' -s 6 -c 2 -o $out/${tag}_q8 -f \
  python bench.py --steps 1 --warmup 3 --frames $frames --no-e2e --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $out/${tag}_q8.ncu-rep
