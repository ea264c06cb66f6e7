#!/bin/bash
cd "$(dirname "$0")/.."
for g in 1 2 4 8; do timeout -s KILL 300 python tools/groups_case.py 4096 $g 2>&1 | tail -2; done
