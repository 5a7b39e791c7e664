#!/bin/bash
# round 2: L2 prefetch distance of the band SLIC kernel
set -u
out=gpurun_out
mkdir -p $out
for n in 8 64 256; do timeout 300 python tools/slic_run.py $n; done 2>&1 | sed "s/^/prefetch 4 rows: /" | tee $out/r2t_slic.txt
for t in 0 8 16; do for n in 8 64 256; do DCMT_LIB=$PWD/depth_completion_mt_b200/variants/libdcmt_slicpf$t.so timeout 300 python tools/slic_run.py $n; done 2>&1 | sed "s/^/prefetch $t rows: /" | tee -a $out/r2t_slic.txt; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_slic_assign_band' -s 4 -c 1 -o $out/r2t_slic_band -f python tools/slic_run.py 64 > $out/r2t_slic_ncu.log 2>&1; echo "slic band ncu rc=$?"
