#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
for v in "" vr208 vr216 vr224 vr232; do
  lib=pymd_b200/libpymd_b200.so; [ -n "$v" ] && lib=tools/libpymd_$v.so
  for B in 4096 2048; do
  PYMD_FUSED3=1 PYMD_B200_LIB=$lib $T 200 python tools/quick_bench.py --mode 2 --batch $B --steps 512 > gpurun_out/c19_${v:-prod}_$B.txt 2>&1
  echo "variant=${v:-prod} B=$B $(head -1 gpurun_out/c19_${v:-prod}_$B.txt | cut -c1-90)"
  done
done
