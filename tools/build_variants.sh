#!/bin/bash
# Variant builds of fused3.cu (register re-balancing / unit offset switches) -> tools/libpymd_v<name>.so, for A/B runs
# on the GPU with PYMD_B200_LIB=...; all other objects come from the production build in pymd_b200/csrc/build.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
B=$ROOT/pymd_b200/csrc/build
build() { # name flags...
  local name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" \
       -c $ROOT/pymd_b200/csrc/fused3.cu -o /tmp/fused3_$name.o 2> /tmp/fused3_$name.log
  grep -A1 "fused3_kernelILi2ELb1ELb1ELi15" /tmp/fused3_$name.log | grep -E "spill|registers" | head -3
  objs=$(ls $B/*.o | grep -v fused3.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/tools/libpymd_v$name.so $objs /tmp/fused3_$name.o -lcudart
  echo built v$name
}
build noreg -DZ_SETMAXNREG=0 &
build nooff -DZ_OFFSET=0 &
build r208 -DZ_ROW_REGS=208 -DZ_BATH_REGS=88 &
build plain -DZ_SETMAXNREG=0 -DZ_OFFSET=0 &
wait
