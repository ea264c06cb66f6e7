#!/bin/bash
# round-2 validation: whole GPU test suite, smoke, default bench, ncu capture of the dominant kernel + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
$T 1800 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.txt 2>&1
tail -4 gpurun_out/final_pytest.txt
$T 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
$T 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/final_bench.json"))
    print("value %.1fM e2e %.1fM frac %.3f whole %.3f"%(d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["frac"], d["roofline"]["whole_step_frac"]), d["e2e"]["breakdown"]["total"], d["shard_check"]["shard_check"], d["timed_seconds"], d["clocks"], d["cpu_baseline"]["value"], d["gpu_launches"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/final_bench.err").read()[-2000:])
PY
$T 600 ncu --set full --import-source on --clock-control none -k regex:fused3_kernel -s 3 -c 1 -o gpurun_out/final_f3 -f python tools/prof_case.py 4096 2 > gpurun_out/final_ncu.log 2>&1
$T 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 3 --warmup 3 --piece 64 --no-e2e --no-cpu --no-shard-check > gpurun_out/final_ncu_bench.log 2>&1
tail -1 gpurun_out/final_ncu.log
