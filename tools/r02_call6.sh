#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
for B in 4096 2048; do
  PYMD_FUSED3=1 $T 200 python tools/quick_bench.py --mode 2 --batch $B --steps 512 > gpurun_out/c6_quick_$B.txt 2>&1
  echo "B=$B $(head -1 gpurun_out/c6_quick_$B.txt | cut -c1-100) $(tail -1 gpurun_out/c6_quick_$B.txt | cut -c60-160)"
  PYMD_FUSED3=1 PYMD_B200_LIB=tools/libpymd_timing.so $T 200 python tools/quick_bench.py --mode 2 --batch $B --steps 256 > gpurun_out/c6_timing_$B.txt 2>&1
  grep timing gpurun_out/c6_timing_$B.txt | tail -2
done
PYMD_FUSED3=1 $T 600 python -m pytest tests/test_gpu_trajectory.py -m gpu -x -q -k "tile_sizes or far_field or randomized" > gpurun_out/c6_pytest.txt 2>&1
tail -2 gpurun_out/c6_pytest.txt
