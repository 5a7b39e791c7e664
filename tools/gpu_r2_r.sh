#!/bin/bash
# round 2: band SLIC with deferred near ties (no call in the walk), thread counts
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_slic.py tests/test_reference_build.py tests/test_cpp_shim.py -q -m gpu > $out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2r_pytest.log
for n in 1 2 4 8 64 256; do timeout 300 python tools/slic_run.py $n; done 2>&1 | tee $out/r2r_slic.txt
DCMT_SLIC_BAND_MIN_FRAMES=1 timeout 300 python tools/slic_run.py 1 2>&1 | sed 's/^/band kernel: /' | tee -a $out/r2r_slic.txt
for t in 1024 896 640; do for n in 8 64 256; do DCMT_LIB=$PWD/depth_completion_mt_b200/variants/libdcmt_slic$t.so timeout 300 python tools/slic_run.py $n; done 2>&1 | sed "s/^/$t threads: /" | tee -a $out/r2r_slic.txt; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_slic_assign_band' -s 4 -c 1 -o $out/r2r_slic_band -f python tools/slic_run.py 64 > $out/r2r_slic_ncu.log 2>&1; echo "slic band ncu rc=$?"
