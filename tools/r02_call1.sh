#!/bin/bash
# round-2 GPU call 1: baseline numbers, fused2 phase timing, new parity tests, sanitizer
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt
python tools/quick_bench.py --mode 2 --batch 4096 --steps 512 > gpurun_out/c1_quick.txt 2>&1
PYMD_B200_LIB=tools/libpymd_timing.so python tools/quick_bench.py --mode 2 --batch 4096 --steps 256 > gpurun_out/c1_timing.txt 2>&1
python tools/quick_bench.py --mode 2 --batch 2048 --steps 512 > gpurun_out/c1_quick2048.txt 2>&1
python tools/e2e_breakdown.py 2048 32768 > gpurun_out/c1_e2e_breakdown.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.txt 2>&1
tail -3 gpurun_out/c1_pytest.txt
bash tools/sanitize.sh > gpurun_out/c1_sanitize.txt 2>&1
cat gpurun_out/c1_quick.txt gpurun_out/c1_timing.txt | tail -8
cat gpurun_out/c1_sanitize.txt | tail -30
