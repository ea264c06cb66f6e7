#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
for cfg in "sweep:64,4000,4096" "sweep:64,1000,4096" "sweep:64,200,4096"; do for mode in 1; do
  $T 600 python bench.py --config $cfg --mode $mode --steps 3 --warmup 3 --no-e2e --no-cpu --no-shard-check > gpurun_out/c18.json 2> gpurun_out/c18.err
  python - "$cfg" $mode <<'PY'
import json, sys
try:
    d=json.load(open("gpurun_out/c18.json"))
    print(sys.argv[1], "mode", sys.argv[2], "%.2f M/s"%(d["value"]/1e6), {k:round(v["share_of_step"],2) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e, open("gpurun_out/c18.err").read()[-800:])
PY
done; done
# 32-step launch of the final fused3 under ncu (traffic, tensor-pipe counter) + launch list of a short bench
PYMD_FUSED3=1 $T 600 ncu --set full --import-source on --clock-control none -k regex:fused3_kernel -s 3 -c 1 -o gpurun_out/c18_f3 -f python tools/prof_case.py 4096 2 > gpurun_out/c18_ncu.log 2>&1
$T 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c18_launches.csv python bench.py --steps 3 --warmup 3 --piece 64 --no-e2e --no-cpu --no-shard-check > gpurun_out/c18_ncu_bench.log 2>&1
tail -2 gpurun_out/c18_ncu.log
