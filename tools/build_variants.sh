#!/bin/bash
# Variant builds of fused3.cu (register re-balancing) -> tools/libpymd_v<name>.so, for A/B runs on the GPU with
# PYMD_B200_LIB=...; all other objects come from the production build in pymd_b200/csrc/build.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
B=$ROOT/pymd_b200/csrc/build
build() { # name flags...
  local name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" \
       -c $ROOT/pymd_b200/csrc/fused3.cu -o /tmp/fused3_$name.o 2> /tmp/fused3_$name.log
  echo "$name: $(grep -A1 'fused3_kernelILi2ELb1ELb1ELi15ELi1' /tmp/fused3_$name.log | grep -E 'spill' | head -1)"
  objs=$(ls $B/*.o | grep -v fused3.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/tools/libpymd_v$name.so $objs /tmp/fused3_$name.o -lcudart
}
build r208 -DZ_ROW_REGS=208 -DZ_BATH_REGS=88 &
build r216 -DZ_ROW_REGS=216 -DZ_BATH_REGS=72 &
build r224 -DZ_ROW_REGS=224 -DZ_BATH_REGS=56 &
build r232 -DZ_ROW_REGS=232 -DZ_BATH_REGS=40 &
wait
