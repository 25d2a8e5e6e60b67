#!/bin/bash
# SASS evidence: opcode histogram of the whole library and per kernel family, plus the DMMA loop of lm_tile_kernel.
LIB=rdis_b200/librdis_b200.so
OUT=profiles/r02_sass_opcodes.txt
{
echo "# cuobjdump -sass $LIB (sm_100a) — opcode histogram, whole library"
cuobjdump -sass $LIB | grep -E "^\s+/\*[0-9a-f]+\*/" | awk '{print $2}' | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn | head -60
echo
echo "# Blackwell / Hopper-class machinery (counts over the whole library)"
for op in UBLKCP UTMALDG LDGSTS SYNCS STAS UCGABAR DMMA HMMA BAR CCTL MUFU DFMA DADD DMUL SHFL; do printf "%-10s %s\n" $op $(cuobjdump -sass $LIB | grep -cE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+\s+)?$op"); done
echo
for k in solve_ba_cameras_kernel solve_ba_points_kernel lm_tile_kernel nlpf_tile_sweep_kernel ba_sweep_kernel solve_strict_kernel; do
  echo "# kernel family $k: opcode histogram (all instantiations)"
  cuobjdump -sass $LIB | awk -v k="$k" '/Function :/ {on = (index($0, k) > 0)} on' | grep -E "^\s+/\*[0-9a-f]+\*/" | awk '{print $2}' | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn | head -16
  echo
done
echo "# lm_tile_kernel<kTrailing>: the DMMA inner loop (first 60 lines around the first DMMA)"
cuobjdump -sass $LIB | awk '/Function :/ {on = (index($0, "lm_tile_kernelILi1E") > 0)} on' | grep -n "DMMA" | head -1
cuobjdump -sass $LIB | awk '/Function :/ {on = (index($0, "lm_tile_kernelILi1E") > 0)} on' | grep -B 12 -A 40 -m 1 "DMMA" | cut -c1-110
} > $OUT
wc -l $OUT
