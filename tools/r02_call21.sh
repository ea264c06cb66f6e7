#!/bin/bash
cd "$(dirname "$0")/.."
for i in 1 2; do
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-shard-check > gpurun_out/c21_$i.json 2> gpurun_out/c21_$i.err
python - $i <<'PY'
import json, sys
d=json.load(open("gpurun_out/c21_%s.json"%sys.argv[1])); e=d["e2e"]
print(e["seconds"], {k:(round(v,4) if isinstance(v,float) else v) for k,v in e["breakdown"].items() if k not in ("note","h2d")})
PY
done
