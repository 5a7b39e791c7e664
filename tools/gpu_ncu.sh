#!/bin/bash
# one ncu --set full capture of the fused kernels (after a plain run of the same command has exited 0)
set -u
tag=${1:-n}
out=gpurun_out
mkdir -p $out
timeout 600 python bench.py --steps 1 --warmup 3 --frames 79 --no-e2e --no-cpu-baseline > $out/${tag}_plain.json 2> $out/${tag}_plain.err; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_q8_(front|tail)' -s 6 -c 2 -o $out/${tag}_q8 -f \
  python bench.py --steps 1 --warmup 3 --frames 79 --no-e2e --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $out/${tag}_q8.ncu-rep
