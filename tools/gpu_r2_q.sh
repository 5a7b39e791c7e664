#!/bin/bash
# round 2: strip-walk band SLIC kernel and the front's own tile width -- parity, then throughput against the previous choices
set -u
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu > $out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2q_pytest.log
for n in 1 2 4 8 64 256; do timeout 300 python tools/slic_run.py $n; done 2>&1 | tee $out/r2q_slic.txt
for n in 1 2 4 8; do DCMT_SLIC_BAND_MIN_FRAMES=100000 timeout 300 python tools/slic_run.py $n; done 2>&1 | sed 's/^/tile kernel: /' | tee -a $out/r2q_slic.txt
for n in 1 64 256; do DCMT_LIB=$PWD/depth_completion_mt_b200/variants/libdcmt_slic1024.so timeout 300 python tools/slic_run.py $n; done 2>&1 | sed 's/^/1024 threads: /' | tee -a $out/r2q_slic.txt
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline'].get('kernels') or {}
print('$2', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), {n: round(v['ms_per_step'],4) for n, v in k.items() if isinstance(v, dict)})"; }
for w in auto 0; do
  if [ $w = auto ]; then unset DCMT_FRONT_TILE_W; else export DCMT_FRONT_TILE_W=$w; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2q_front_$w.json 2>> $out/r2q.err; show $out/r2q_front_$w.json "front width $w 352x1216:"
  timeout 300 python bench.py --input u16 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2q_front_u16_$w.json 2>> $out/r2q.err; show $out/r2q_front_u16_$w.json "front width $w 352x1216 u16:"
  for shape in "375 1242 1024" "512 1760 512" "1024 2048 192" "2048 4096 48"; do
    set -- $shape
    timeout 300 python bench.py --rows $1 --cols $2 --frames $3 --density 0.05 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2q_front_${w}_$1.json 2>> $out/r2q.err; show $out/r2q_front_${w}_$1.json "front width $w $1x$2:"
  done
done
unset DCMT_FRONT_TILE_W
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/r2q_launches_slic.csv python tools/slic_run.py 64 > $out/r2q_ncu1.log 2>&1
python tools/summarize_launches.py $out/r2q_launches_slic.csv 2>/dev/null | head -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_slic_assign_band' -s 4 -c 1 -o $out/r2q_slic_band -f python tools/slic_run.py 64 > $out/r2q_slic_ncu.log 2>&1; echo "slic band ncu rc=$?"
