#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
$T 900 python -m pytest tests/test_gpu_trajectory.py -m gpu -x -q -k "tile_sizes or far_field or randomized or short_horizon or switching or size_independent or baseline_config or golden" > gpurun_out/c4_pytest.txt 2>&1
tail -5 gpurun_out/c4_pytest.txt
for v in "" vnoreg vnooff vr208 vplain; do for B in 4096 2048; do
  lib=pymd_b200/libpymd_b200.so; [ -n "$v" ] && lib=tools/libpymd_$v.so
  PYMD_B200_LIB=$lib $T 200 python tools/quick_bench.py --mode 2 --batch $B --steps 512 > gpurun_out/c4_quick_${v:-prod}_$B.txt 2>&1
  echo "variant=${v:-prod} B=$B $(tail -1 gpurun_out/c4_quick_${v:-prod}_$B.txt | cut -c1-160)"
  head -1 gpurun_out/c4_quick_${v:-prod}_$B.txt | cut -c1-120
done; done
