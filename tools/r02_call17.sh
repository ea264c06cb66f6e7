#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
$T 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/c17_bench2.json 2> gpurun_out/c17_bench2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c17_bench2.json").read().strip().splitlines()[-1])
    print("N=2 value %.1fM e2e %.1fM"%(d["value"]/1e6, d["e2e"]["value"]/1e6), d["shard_check"], d["e2e"]["breakdown"]["total"], d["e2e"]["mean_current_nW"])
except Exception as e:
    print("N=2 failed", e); print(open("gpurun_out/c17_bench2.err").read()[-3000:])
PY
$T 600 python bench.py --impl reference --steps 6 --warmup 3 > gpurun_out/c17_ref.json 2> gpurun_out/c17_ref.err; cut -c1-400 gpurun_out/c17_ref.json
