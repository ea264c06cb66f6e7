#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
PYMD_FUSED3=1 $T 600 ncu --set full --import-source on --clock-control none -k regex:fused3_kernel -s 4 -c 1 -o gpurun_out/c14_f3_4096 -f python tools/prof_case.py 4096 2 > gpurun_out/c14_ncu1.log 2>&1
PYMD_FUSED3=1 $T 600 ncu --set full --import-source on --clock-control none -k regex:fused3_kernel -s 4 -c 1 -o gpurun_out/c14_f3_2048 -f python tools/prof_case.py 2048 2 > gpurun_out/c14_ncu2.log 2>&1
tail -1 gpurun_out/c14_ncu1.log gpurun_out/c14_ncu2.log
