#!/bin/bash
# GPU box: parity tests (incl. GPU vs the reference build), smoke(), the default bench line and the reference CPU arm.
set -u
out=gpurun_out; mkdir -p $out
ls -la oracle/_ref/ > $out/refcheck_ls.txt 2>&1
timeout 1700 python -m pytest tests -x -q -m gpu > $out/refcheck_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/refcheck_pytest_gpu.log
tail -4 $out/refcheck_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/refcheck_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/refcheck_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/refcheck_bench_reference.json 2> $out/refcheck_bench.err; echo "ref rc=$?"
cat $out/refcheck_bench_reference.json
timeout 900 python bench.py > $out/refcheck_bench.json 2>> $out/refcheck_bench.err; echo "bench rc=$?"
cat $out/refcheck_bench.json
