#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
for B in 4096 2048; do
  PYMD_FUSED3=1 $T 200 python tools/quick_bench.py --mode 2 --batch $B --steps 512 > gpurun_out/chk5_$B.txt 2>&1
  echo "f3 B=$B $(head -1 gpurun_out/chk5_$B.txt | cut -c1-100)"
done
PYMD_FUSED3=1 $T 200 python tools/quick_bench.py --mode 2 --batch 4096 --steps 512 --case chain39 > gpurun_out/chk5_c39.txt 2>&1
echo "chain39 $(head -1 gpurun_out/chk5_c39.txt | cut -c1-100) $(tail -1 gpurun_out/chk5_c39.txt | cut -c60-160)"
$T 900 python -m pytest tests/test_gpu_trajectory.py tests/test_gpu_run.py -m gpu -x -q -k "tile_sizes or far_field or randomized or short_horizon or size_independent or baseline_config or overlaid or pipelined" > gpurun_out/chk_pytest.txt 2>&1
tail -3 gpurun_out/chk_pytest.txt
