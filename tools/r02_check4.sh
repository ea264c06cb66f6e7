#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/legs_launches.csv python tools/legs_case.py 4096 32768 > gpurun_out/legs.log 2>&1
tail -2 gpurun_out/legs.log
python tools/launch_list.py gpurun_out/legs_launches.csv 400 | grep "^#"
