#!/bin/bash
# One GPU-box session: parity tests, bench lines, ncu launch list and one full capture of the fused kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|notests]
set -u
tag=${1:-run}
mode=${2:-tests}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.txt 2>&1
nproc > $out/${tag}_cpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $out/${tag}_cpu.txt
if [ "$mode" = tests ]; then
  timeout 1500 python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
  tail -3 $out/${tag}_pytest_gpu.log
fi
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
cat $out/${tag}_bench.json
timeout 600 python bench.py --input u16 --no-cpu-baseline > $out/${tag}_bench_u16.json 2>> $out/${tag}_bench.err; echo "bench u16 rc=$?"
cat $out/${tag}_bench_u16.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --frames 158 --no-e2e --no-cpu-baseline > $out/${tag}_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_q8_ -s 18 -c 4 -o $out/${tag}_q8 -f \
  python bench.py --steps 1 --warmup 3 --frames 79 --no-e2e --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $out | tail -20
