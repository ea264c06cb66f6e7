#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "powerspec" 2>&1 | tail -5
timeout -s KILL 300 python tools/spec_case.py 2>&1 | tail -4
timeout -s KILL 300 python tools/spec_case.py 16384 60 1024 2>&1 | tail -4
