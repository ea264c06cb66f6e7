#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
$T 900 python -m pytest tests/test_gpu_run.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/chk_pytest.txt 2>&1
tail -12 gpurun_out/chk_pytest.txt
for i in 1 2; do
$T 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-shard-check > gpurun_out/chk_bench_$i.json 2> gpurun_out/chk_bench_$i.err
python - $i <<'PY'
import json, sys
try:
    d=json.load(open("gpurun_out/chk_bench_%s.json"%sys.argv[1])); e=d["e2e"]
    print("e2e %.1fM"%(e["value"]/1e6), e["seconds"], e["sub_batch"], e.get("record_overlay"), {k:(round(v,4) if isinstance(v,float) else v) for k,v in e["breakdown"].items() if k not in ("note","h2d")}, e["mean_current_nW"])
except Exception as ex:
    print("bench failed", ex); print(open("gpurun_out/chk_bench_%s.err"%sys.argv[1]).read()[-2500:])
PY
done
$T 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-shard-check --no-record-overlay > gpurun_out/chk_bench_n.json 2> gpurun_out/chk_bench_n.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/chk_bench_n.json")); e=d["e2e"]
print("no-overlay e2e %.1fM"%(e["value"]/1e6), e["seconds"], e["sub_batch"], e["mean_current_nW"])
PY
