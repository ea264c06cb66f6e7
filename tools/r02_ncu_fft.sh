#!/bin/bash
# ncu captures of the round-2 FFT kernels: fused last spectrum pass and the cp.async-staged radix-128 pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:fft_pass128_power_kernel -s 2 -c 1 -o gpurun_out/ncu_power -f python tools/spec_case.py 32768 60 512 > gpurun_out/ncu_power.log 2>&1
timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:fft_pass128s_kernel -s 2 -c 1 -o gpurun_out/ncu_pass128s -f python tools/fft_case.py 16384 21120 > gpurun_out/ncu_pass128s.log 2>&1
tail -1 gpurun_out/ncu_power.log gpurun_out/ncu_pass128s.log
exit 0
