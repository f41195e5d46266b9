#!/bin/bash
# Opcode evidence per kernel of the in-tree libflinkms_b200.so (cuobjdump -sass): which Blackwell paths the product really uses.
#   tools/sass_summary.sh > profiles/r02_sass_summary.txt
SO=${1:-flink-ms_b200/libflinkms_b200.so}
echo "# cuobjdump -sass $SO  ($(date -u +%Y-%m-%dT%H:%MZ), $(nvcc --version | tail -1))"
echo "# mnemonics: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = TMA tensor load (gather4), UBLKCP = cp.async.bulk,"
echo "#            LDGSTS = cp.async, HMMA = mma.sync (warp-level tensor path), SYNCS = mbarrier, FFMA2 = packed fp32 FMA, MUFU = rsqrt"
echo "# kernel | opcode | static count      (TOTAL = all SASS instructions of the kernel)"
cuobjdump -sass "$SO" | awk '
/Function :/ { fn=$3; next }
/^[ \t]+\/\*[0-9a-f]+\*\// {
  op=$2; if (op ~ /^@/) op=$3; split(op, a, "."); m=a[1];
  if (m=="UTCHMMA"||m=="LDTM"||m=="UTCBAR"||m=="UTMALDG"||m=="UBLKCP"||m=="LDGSTS"||m=="HMMA"||m=="SYNCS"||m=="FFMA2"||m=="MUFU"||m=="STTM") c[fn"\t"op]++;
  n[fn]++
}
END { for (k in c) print k "\t" c[k]; for (f in n) print f "\tTOTAL\t" n[f] }' | sort | while IFS=$'\t' read -r fn op cnt; do
  short=$(echo "$fn" | c++filt | sed -E 's/^void //; s/\(.*$//')
  printf "%-70s %-28s %6s\n" "$short" "$op" "$cnt"
done
