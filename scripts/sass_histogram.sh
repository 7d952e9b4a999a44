#!/bin/bash
# Opcode evidence for the Blackwell-native paths of libsimclr_b200.so (run anywhere: cuobjdump needs no GPU).
# Usage: bash scripts/sass_histogram.sh > profiles/rNN_sass_histogram.txt
cd "$(dirname "$0")/.."
SO=simclr_b200/libsimclr_b200.so
TMP=$(mktemp)
cuobjdump -sass "$SO" > "$TMP" 2>/dev/null
echo "# SASS opcode counts of $SO ($(git rev-parse --short HEAD 2>/dev/null)), cuobjdump -sass, arch $(grep -m1 'arch =' "$TMP" | sed 's/.*= //')"
echo "# tcgen05.mma -> UTCHMMA; tcgen05.ld -> LDTM; tcgen05.commit -> UTCBAR; TMA tile / im2col loads -> UTMALDG; TMA stores / reduce-add -> UTMASTG / UTMAREDG"
for op in UTCHMMA LDTM UTCBAR "UTMALDG\.2D" "UTMALDG\.4D\.IM2COL" "UTMALDG\.4D " "UTMASTG\.2D" "UTMASTG\.4D" "UTMAREDG" "SYNCS\." "LDGSTS" "ELECT" "FFMA2" "HMMA\." ; do
  printf "%-24s %6d\n" "$(echo $op | sed 's/\\//g')" "$(grep -c -E "\b$op" "$TMP")"
done
echo
echo "# per kernel family: UTCHMMA / LDTM / UTMALDG / UTMASTG+UTMAREDG"
awk '/Function : /{name=$3} /UTCHMMA/{m[name]++} /LDTM/{l[name]++} /UTMALDG/{t[name]++} /UTMASTG|UTMAREDG/{s[name]++} END{for(k in m) printf "%5d %5d %5d %5d  %s\n", m[k], l[k], t[k], s[k], k}' "$TMP" | c++filt | sed 's/(CUtensorMap_st.*//' | sort -k5 | awk '{a=$5; for(i=6;i<=NF;i++)a=a" "$i; print $1, $2, $3, $4, substr(a,1,140)}'
rm -f "$TMP"
