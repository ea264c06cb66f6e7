#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
for f3 in 0 1; do
PYMD_FUSED3=$f3 $T 600 ncu --set full --import-source on --clock-control none -k regex:fused[23]_kernel -s 4 -c 1 -o gpurun_out/c3_f3$f3 -f python tools/prof_case.py 4096 2 > gpurun_out/c3_ncu_f3$f3.log 2>&1
tail -2 gpurun_out/c3_ncu_f3$f3.log
done
ls -la gpurun_out/c3_*
