#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
$T 900 python -m pytest tests/test_gpu_trajectory.py -m gpu -x -q -k "tile_sizes or far_field or randomized or short_horizon or switching or size_independent or baseline_config or golden" > gpurun_out/chk_pytest.txt 2>&1
tail -3 gpurun_out/chk_pytest.txt
POINTS="sweep:64,200,4096 sweep:60,200,65536,15 cfg4 cfg2" bash tools/sweep.sh 1 | cut -c1-200
