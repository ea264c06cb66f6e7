#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
$T 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c8_pytest.txt 2>&1
tail -4 gpurun_out/c8_pytest.txt
$T 900 python bench.py --steps 8 --warmup 3 > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/c8_bench.json"))
    print("value %.1fM e2e %.1fM sub %d"%(d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["sub_batch"]), d["e2e"]["breakdown"], d["shard_check"]["shard_check"], d["timed_seconds"], d["clocks"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/c8_bench.err").read()[-2000:])
PY
