#!/bin/bash
# BASELINE.json config 5: synthetic scaling sweep through the driver-visible bench (bench.py --config sweep:nd,ml,B[,nc]),
# one JSON line per point -> gpurun_out/sweep.jsonl (copy to profiles/r02_sweep.jsonl).  Leads of nd/4 dofs each
# unless a fourth field gives nc.  Usage: tools/sweep.sh [gpus] < list of points, or with the built-in list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
G=${1:-1}
OUT=gpurun_out/sweep_${G}gpu.jsonl
: > $OUT
POINTS=${POINTS:-"sweep:64,200,1024 sweep:64,200,4096 sweep:64,200,16384 sweep:64,200,65536 sweep:64,1000,4096 sweep:64,4000,4096 \
sweep:128,200,4096 sweep:128,1000,4096 sweep:128,4000,1024 \
sweep:256,200,1024 sweep:256,200,4096 sweep:256,200,16384 sweep:256,1000,4096 sweep:256,4000,1024 \
sweep:512,200,4096 sweep:512,1000,4096 sweep:512,4000,1024 \
sweep:1024,200,4096 sweep:1024,1000,1024 \
sweep:2048,200,1024 sweep:2048,1000,1024 \
sweep:60,200,65536,15 sweep:255,1000,4096,15 cfg4s cfg4 cfg3"}
for p in $POINTS; do
  extra=""
  [ "$p" = "cfg3" ] && [ "$G" = "1" ] && extra="--batch 1024"     # what each GPU of the 8-GPU layout does
  if [ "$G" = "1" ]; then
    timeout -s KILL 900 python bench.py --config $p --steps 3 --warmup 3 --no-e2e --no-cpu --no-shard-check $extra > gpurun_out/sweep_point.json 2> gpurun_out/sweep_point.err
  else
    timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --config $p --steps 3 --warmup 3 --no-e2e --no-cpu --no-shard-check > gpurun_out/sweep_point.json 2> gpurun_out/sweep_point.err
  fi
  if [ -s gpurun_out/sweep_point.json ]; then
    python - "$p" >> $OUT <<'PY'
import json, sys
d = json.loads(open("gpurun_out/sweep_point.json").read().strip().splitlines()[-1])
r = d["roofline"]
out = dict(config=d["config"]["config"], workload=d["config"]["workload"], n_gpus=d["n_gpus"], batch_per_gpu=d["config"]["batch_per_gpu"],
           traj_steps_per_s=d["value"], ms_per_md_step=d["ms_per_step"] / d["config"]["md_steps_per_step"], setup_s=d["setup_s"],
           whole_step_frac_of_fp64_peak=r["whole_step_frac"], whole_step_frac_of_hbm=r["whole_step_frac_of_hbm"],
           flops_per_trajectory_step=r["flops_per_trajectory_step"]["total"], bytes_per_trajectory_step=r["flops_per_trajectory_step"]["bytes"],
           fdl=r["flops_per_trajectory_step"].get("fdl"),
           kernels={k: dict(share=round(v["share_of_step"], 3), us_per_launch=round(v["us_per_launch"], 1), frac=v.get("frac")) for k, v in r["kernels"].items()},
           finite=d["finite"], clocks=d["clocks"])
print(json.dumps(out))
PY
    tail -1 $OUT | cut -c1-230
  else
    echo "{\"config\": \"$p\", \"failed\": \"$(tail -3 gpurun_out/sweep_point.err | tr '\n"' ' _' | cut -c1-300)\"}" >> $OUT
    tail -1 $OUT
  fi
done
