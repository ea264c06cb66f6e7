#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
$T 900 python -m pytest tests/test_gpu_trajectory.py -m gpu -x -q -k "delay_line or large_junctions or switching or several_constraint or short_horizon or golden" > gpurun_out/c10_pytest.txt 2>&1
tail -4 gpurun_out/c10_pytest.txt
run() { # name args
  local name=$1; shift
  $T 900 python bench.py "$@" --no-cpu > gpurun_out/c10_$name.json 2> gpurun_out/c10_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c10_$name.json"))
    e=d.get("e2e") or {}
    print("$name value %.3fM e2e %.3fM ms/step %.1f piece %d setup %.1f"%(d["value"]/1e6, (e.get("value") or 0)/1e6, d["ms_per_step"], d["config"]["md_steps_per_step"], d["setup_s"]), {k:(round(v["share_of_step"],3), round(v["us_per_launch"],1), round(v["launches_per_md_step"],2)) for k,v in d["roofline"]["kernels"].items()}, "whole", round(d["roofline"]["whole_step_frac"],3))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/c10_$name.err").read()[-1500:])
PY
}
run cfg4s --config cfg4s --steps 3 --warmup 3 --no-e2e
run nd129 --config sweep:129,200,4096,30 --steps 3 --warmup 3 --no-e2e
run cfg3 --config cfg3 --steps 5 --warmup 3 --no-e2e
run cfg4 --config cfg4 --steps 5 --warmup 3 --no-e2e
run cfg2 --steps 8 --warmup 3
