#!/bin/bash
# Variant builds of fused.cu (fused2 register re-balancing) -> tools/libpymd_f2<name>.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
B=$ROOT/pymd_b200/csrc/build
build() {
  local name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" \
       -c $ROOT/pymd_b200/csrc/fused.cu -o /tmp/fused_$name.o 2> /tmp/fused_$name.log
  grep -A1 "fused2_kernelILi[24]ELb1ELb1ELi3" /tmp/fused_$name.log | grep -E "spill" | head -3
  objs=$(ls $B/*.o | grep -v "/fused.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/tools/libpymd_f2$name.so $objs /tmp/fused_$name.o -lcudart
  echo built f2$name
}
build r192 -DF2_SETMAXNREG=1 -DF2_ROW_REGS=192 -DF2_BATH_REGS=120 &
build r200 -DF2_SETMAXNREG=1 -DF2_ROW_REGS=200 -DF2_BATH_REGS=104 &
wait
