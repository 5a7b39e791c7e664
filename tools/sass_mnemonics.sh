#!/bin/bash
# static SASS mnemonic counts of the fused kernels (run where the library was built; no GPU needed)
lib=${1:-depth_completion_mt_b200/libdcmt.so}
echo "# static SASS mnemonic counts of the fused kernels in libdcmt.so (cuobjdump -sass, sm_100a):"
echo "# UTMALDG = TMA tensor load, SYNCS = mbarrier, VIMNMX(3).U16x2 = packed 16-bit min/max (two comparisons per VIMNMX3), IDP.2A = 16x8-bit dot product"
for k in "k_q8_frontILb0" "k_q8_guided_frontILb1" "k_q8_tail"; do
  echo "== $k"
  cuobjdump -sass "$lib" | awk -v k="$k" '/Function :/{f=index($0,k)>0} f' | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" | awk '{print $NF}' | sort | uniq -c | sort -rn | head -24
done
echo "== k_q8_tail: TMA / mbarrier instructions (all occurrences)"
cuobjdump -sass "$lib" | awk '/Function :/{f=index($0,"k_q8_tail")>0} f' | grep -E "UTMALDG|SYNCS"
