#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
$T 900 python -m pytest tests/test_gpu_trajectory.py -m gpu -x -q -k "tile_sizes or far_field or randomized or short_horizon or switching or size_independent or baseline_config" > gpurun_out/c2_pytest.txt 2>&1
tail -5 gpurun_out/c2_pytest.txt
for f3 in 1 0; do for B in 4096 2048; do
  PYMD_FUSED3=$f3 $T 200 python tools/quick_bench.py --mode 2 --batch $B --steps 512 > gpurun_out/c2_quick_f3${f3}_$B.txt 2>&1
  echo "f3=$f3 B=$B"; tail -2 gpurun_out/c2_quick_f3${f3}_$B.txt
done; done
PYMD_B200_LIB=tools/libpymd_timing.so $T 200 python tools/quick_bench.py --mode 2 --batch 4096 --steps 256 > gpurun_out/c2_timing.txt 2>&1
grep timing gpurun_out/c2_timing.txt | tail -2
PYMD_B200_LIB=tools/libpymd_timing.so $T 200 python tools/quick_bench.py --mode 2 --batch 2048 --steps 256 > gpurun_out/c2_timing2048.txt 2>&1
grep timing gpurun_out/c2_timing2048.txt | tail -2
