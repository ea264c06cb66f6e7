#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="timeout -s KILL"
$T 200 python tools/quick_bench.py --mode 2 --batch 4096 --steps 512 > gpurun_out/c7_quick_4096.txt 2>&1
echo "f3 B=4096 $(head -1 gpurun_out/c7_quick_4096.txt | cut -c1-100) $(tail -1 gpurun_out/c7_quick_4096.txt | cut -c60-160)"
$T 600 python -m pytest tests/test_gpu_run.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/c7_pytest.txt 2>&1
tail -3 gpurun_out/c7_pytest.txt
$T 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c7_bench_overlap.json 2> gpurun_out/c7_bench_overlap.err
tail -c 3000 gpurun_out/c7_bench_overlap.json; tail -5 gpurun_out/c7_bench_overlap.err
$T 900 python bench.py --steps 5 --warmup 3 --no-overlap --no-cpu --no-shard-check > gpurun_out/c7_bench_serial.json 2> gpurun_out/c7_bench_serial.err
python - <<'PY'
import json
for f in ("overlap","serial"):
    try:
        d=json.load(open("gpurun_out/c7_bench_%s.json"%f))
        print(f, "value %.1fM e2e %.1fM"%(d["value"]/1e6, d["e2e"]["value"]/1e6), d["e2e"]["breakdown"])
    except Exception as e: print(f, "failed", e)
PY
