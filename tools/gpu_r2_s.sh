#!/bin/bash
# round 2: band SLIC with a queue of deferred pixels; two CTAs of 384 threads per SM (default) against one of 768 (with / without the
# double-precision centres in shared memory) and two of 256
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_slic.py tests/test_reference_build.py tests/test_cpp_shim.py -q -m gpu > $out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2s_pytest.log
for n in 1 2 4 8 64 256; do timeout 300 python tools/slic_run.py $n; done 2>&1 | tee $out/r2s_slic.txt
DCMT_SLIC_BAND_MIN_FRAMES=1 timeout 300 python tools/slic_run.py 1 2>&1 | sed 's/^/band kernel: /' | tee -a $out/r2s_slic.txt
for t in 768c 768 256; do for n in 8 64 256; do DCMT_LIB=$PWD/depth_completion_mt_b200/variants/libdcmt_slic$t.so timeout 300 python tools/slic_run.py $n; done 2>&1 | sed "s/^/$t: /" | tee -a $out/r2s_slic.txt; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_slic_assign_band' -s 4 -c 1 -o $out/r2s_slic_band -f python tools/slic_run.py 64 > $out/r2s_slic_ncu.log 2>&1; echo "slic band ncu rc=$?"
