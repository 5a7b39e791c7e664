#!/bin/bash
# round 2: the float-walk band SLIC kernel (parity, throughput by batch size, against the tile kernel and a 1024-thread build),
# and other tile widths for k_q8_front
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_slic.py tests/test_stereo_gray.py tests/test_reference_build.py tests/test_cpp_shim.py -q -m gpu > $out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2p_pytest.log
for n in 1 8 64 256; do timeout 300 python tools/slic_run.py $n; done 2>&1 | tee $out/r2p_slic.txt
for n in 1 64; do DCMT_SLIC_BAND_MIN_FRAMES=100000 timeout 300 python tools/slic_run.py $n; done 2>&1 | sed 's/^/tile kernel: /' | tee -a $out/r2p_slic.txt
for n in 1 64 256; do DCMT_LIB=$PWD/depth_completion_mt_b200/variants/libdcmt_slic1024.so timeout 300 python tools/slic_run.py $n; done 2>&1 | sed 's/^/1024 threads: /' | tee -a $out/r2p_slic.txt
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline'].get('kernels')
print('$2', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'kernels', k)"; }
for w in 0 112 176; do
  DCMT_FRONT_TILE_W=$w timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2p_front_w$w.json 2>> $out/r2p.err; show $out/r2p_front_w$w.json "front tile width $w:"
done
DCMT_LIB=$PWD/depth_completion_mt_b200/variants/libdcmt_front512.so DCMT_FRONT_TILE_W=304 DCMT_FRONT_TILE_H=44 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2p_front_512.json 2>> $out/r2p.err; show $out/r2p_front_512.json "front 2 x 512 threads, 304 x 44 tiles:"
DCMT_LIB=$PWD/depth_completion_mt_b200/variants/libdcmt_front512.so timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/r2p_front_512b.json 2>> $out/r2p.err; show $out/r2p_front_512b.json "front 2 x 512 threads, tail's tiles:"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/r2p_launches_slic.csv python tools/slic_run.py 64 > $out/r2p_ncu1.log 2>&1
python tools/summarize_launches.py $out/r2p_launches_slic.csv 2>/dev/null | head -10
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_slic_assign_band' -s 4 -c 1 -o $out/r2p_slic_band -f python tools/slic_run.py 64 > $out/r2p_slic_ncu.log 2>&1; echo "slic band ncu rc=$?"
