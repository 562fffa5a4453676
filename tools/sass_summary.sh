#!/bin/bash
# Blackwell-native evidence in the tree: per kernel of libodevit.so, how many tcgen05 MMA (UTCHMMA), tensor-memory load /
# store (LDTM / STTM), TMA tensor load / store (UTMALDG / UTMASTG) and bulk-copy (UBLKCP) instructions its SASS holds.
#   tools/sass_summary.sh > profiles/rNN_sass_summary.txt
LIB=${1:-odevit_b200/csrc/libodevit.so}
echo "# $(basename $LIB): $(cuobjdump -sass $LIB | grep -c 'Function :') kernels, arch $(cuobjdump -lelf $LIB | head -1)"
cuobjdump -sass "$LIB" | awk '
/Function :/ { name=$3; order[++n]=name }
/UTCHMMA/ { mma[name]++; if ($0 ~ /2CTA/) mma2[name]++ }
/LDTM/ { ldtm[name]++ }  /STTM/ { sttm[name]++ }  /UTMALDG/ { tmal[name]++ }  /UTMASTG/ { tmas[name]++ }  /UBLKCP/ { blk[name]++ }
END {
  printf "%-8s %-8s %-6s %-6s %-8s %-8s %-7s kernel\n", "UTCHMMA", "(.2CTA)", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP";
  for (i=1;i<=n;i++) { k=order[i]; if (mma[k]+ldtm[k]+sttm[k]+tmal[k]+tmas[k]+blk[k] > 0) {
    printf "%-8d %-8d %-6d %-6d %-8d %-8d %-7d %s\n", mma[k], mma2[k], ldtm[k], sttm[k], tmal[k], tmas[k], blk[k], substr(k,1,150);
    T1+=mma[k]; T2+=mma2[k]; T3+=ldtm[k]; T4+=sttm[k]; T5+=tmal[k]; T6+=tmas[k]; T7+=blk[k] } }
  printf "%-8d %-8d %-6d %-6d %-8d %-8d %-7d TOTAL\n", T1, T2, T3, T4, T5, T6, T7 }' | c++filt | sed -e 's/odevit::(anonymous namespace):://' -e 's/(CUtensorMap_st.*//' | cut -c1-170
