#!/bin/bash
# compute-sanitizer racecheck + memcheck of the step kernels (ring buffers, named barriers, staging buffers
# reused across phases).  Summaries -> gpurun_out/sanitizer_*.txt (copy the tails into profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name tool env... -- command
  local name=$1 tool=$2; shift 2
  echo "== $name ($tool): $*" | tee gpurun_out/sanitizer_${name}_${tool}.txt
  timeout 900 env "$@" >> gpurun_out/sanitizer_${name}_${tool}.txt 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_${name}_${tool}.txt
  grep -E "RACECHECK SUMMARY|ERROR SUMMARY|san_case ok|smoke:|rc=" gpurun_out/sanitizer_${name}_${tool}.txt | tail -4
}
S="compute-sanitizer --print-limit 5"
run smoke racecheck   X=1 $S --tool racecheck python tests/smoke_case.py
run smoke memcheck    X=1 $S --tool memcheck python tests/smoke_case.py
run fused2 racecheck  PYMD_FUSED_NT=4 $S --tool racecheck python tools/san_case.py aubz 40 2 40
run fused2 memcheck   PYMD_FUSED_NT=4 $S --tool memcheck python tools/san_case.py aubz 40 2 40
run fused8 racecheck  PYMD_FUSED_NT=2 PYMD_FUSED2=0 $S --tool racecheck python tools/san_case.py aubz 40 2 40
run local racecheck   X=1 $S --tool racecheck python tools/san_case.py alkane 24 0 24
run local memcheck    X=1 $S --tool memcheck python tools/san_case.py alkane 24 0 24
run pipeline memcheck X=1 $S --tool memcheck python tools/san_case.py nd129 34 0 12
