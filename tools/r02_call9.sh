#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
$T 900 python -m pytest tests/test_gpu_trajectory.py -m gpu -x -q -k "delay_line or large_junctions or switching or several_constraint or short_horizon" > gpurun_out/c9_pytest.txt 2>&1
tail -15 gpurun_out/c9_pytest.txt
for f in 0 1; do
PYMD_FDL=$f $T 600 python bench.py --config cfg4s --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/c9_cfg4s_fdl$f.json 2> gpurun_out/c9_cfg4s_fdl$f.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c9_cfg4s_fdl$f.json"))
    print("cfg4s fdl=$f value %.3fM ms/step %.1f piece %d setup %.1f"%(d["value"]/1e6, d["ms_per_step"], d["config"]["md_steps_per_step"], d["setup_s"]), {k:(round(v["share_of_step"],3), round(v["us_per_launch"],1), round(v["launches_per_md_step"],2)) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("cfg4s fdl=$f failed", e); print(open("gpurun_out/c9_cfg4s_fdl$f.err").read()[-1500:])
PY
done
