#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "noise" 2>&1 | tail -3
timeout -s KILL 300 python tools/noise_case.py 4096 2>&1 | tail -3
