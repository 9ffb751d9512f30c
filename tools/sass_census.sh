#!/bin/bash
# Opcode census of the built library: the tcgen05 / TMEM / TMA / cluster opcodes of the SYRK kernels and the fp64 / popc
# mix of the sweep kernel, from `cuobjdump -sass` of the in-tree .so.   tools/sass_census.sh > profiles/r2_sass_census.txt
set -e
cd "$(dirname "$0")/.."
SO=visreps_b200/lib/libvisreps_b200.so
echo "# cuobjdump -sass $SO | opcode counts (whole library), $(date -u +%F)"
cuobjdump -sass $SO > /tmp/vr_all.sass
for op in UTCHMMA UTCHMMA.2CTA LDTM UTMALDG UTCBAR SYNCS UTMACCTL DFMA DADD POPC BMSK IMAD.WIDE LDGSTS "LDG.E.64" "LDS.128" REDG ATOMG; do
  printf "%-14s %6d\n" "$op" "$(grep -c -- " $op" /tmp/vr_all.sass || true)"
done
echo
echo "# per kernel: instructions / UTCHMMA / LDTM / UTMALDG / DFMA / POPC"
awk '/Function :/{name=$3} /^ +\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\//{n[name]++; if ($0 ~ /UTCHMMA/) h[name]++; if ($0 ~ /LDTM/) l[name]++; if ($0 ~ /UTMALDG/) t[name]++; if ($0 ~ / DFMA/) d[name]++; if ($0 ~ / POPC/) p[name]++}
     END{for (k in n) if (k ~ /syrk|sweep_kernelILi8ELi8|sweep_kernelILi1ELi8ELb1|count_rows|count_totals|tie_sums|row_center/) printf "%-90s %6d %4d %4d %4d %4d %4d\n", substr(k,1,90), n[k], h[k], l[k], t[k], d[k], p[k]}' /tmp/vr_all.sass | sort
