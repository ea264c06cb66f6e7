#!/bin/bash
# Debug build with per-phase clock64 instrumentation of the fused kernel -> tools/libpymd_timing.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=/tmp/pymd_timing_build
mkdir -p $OUT
for f in engine gemm fused fused3 local far fdl fft noise spectrum init setup; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DPYMD_FUSED_TIMING \
       -c $ROOT/pymd_b200/csrc/$f.cu -o $OUT/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/tools/libpymd_timing.so $OUT/*.o -lcudart
echo built $ROOT/tools/libpymd_timing.so
