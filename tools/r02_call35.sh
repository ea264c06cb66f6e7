#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -5
for s in 0 1; do
  echo "PYMD_FFT_STAGED=$s"
  PYMD_FFT_STAGED=$s timeout -s KILL 300 python tools/fft_case.py 16384 21120 2>&1 | tail -1
  PYMD_FFT_STAGED=$s timeout -s KILL 300 python tools/fft_case.py 16384 61440 2>&1 | tail -1
  PYMD_FFT_STAGED=$s timeout -s KILL 300 python tools/spec_case.py 2>&1 | tail -3
done
