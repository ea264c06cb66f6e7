#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
for v in f2r192 f2r200; do for B in 4096 2048; do
  PYMD_FUSED3=0 PYMD_B200_LIB=tools/libpymd_$v.so $T 200 python tools/quick_bench.py --mode 2 --batch $B --steps 512 > gpurun_out/c5_quick_${v}_$B.txt 2>&1
  echo "variant=$v B=$B $(head -1 gpurun_out/c5_quick_${v}_$B.txt | cut -c1-100) $(tail -1 gpurun_out/c5_quick_${v}_$B.txt | cut -c60-160)"
done; done
PYMD_B200_LIB=tools/libpymd_timing.so $T 200 python tools/quick_bench.py --mode 2 --batch 4096 --steps 256 > gpurun_out/c5_timing.txt 2>&1
grep timing gpurun_out/c5_timing.txt | tail -2
PYMD_B200_LIB=tools/libpymd_timing.so $T 200 python tools/quick_bench.py --mode 2 --batch 2048 --steps 256 > gpurun_out/c5_timing2048.txt 2>&1
grep timing gpurun_out/c5_timing2048.txt | tail -2
$T 600 ncu --set full --import-source on --clock-control none -k regex:fused3_kernel -s 4 -c 1 -o gpurun_out/c5_f3 -f python tools/prof_case.py 4096 2 > gpurun_out/c5_ncu.log 2>&1
tail -1 gpurun_out/c5_ncu.log
