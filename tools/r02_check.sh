#!/bin/bash
# quick check of the per-step pipeline after a change: its parity tests and three throughput points
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
$T 900 python -m pytest tests/test_gpu_trajectory.py tests/test_gpu_run.py -m gpu -x -q -k "delay_line or large_junctions or switching or several_constraint or short_horizon or golden or pipelined or operands or resumed or fhis" > gpurun_out/chk_pytest.txt 2>&1
tail -3 gpurun_out/chk_pytest.txt
POINTS="sweep:128,200,4096 sweep:256,200,4096 sweep:255,1000,4096,15 cfg4s" bash tools/sweep.sh 1 | cut -c1-200
