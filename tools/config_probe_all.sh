#!/bin/bash
# The other BASELINE configurations on one GPU (tools/config_probe.py), one JSON line each.
P="python $(dirname "$0")/config_probe.py"
$P --label cfg3_alkane_8192 --nd 96 --ncl 0 --batch 8192 --nmd 2048 --steps 1024
$P --label cfg3_alkane_1024_per_gpu --nd 96 --ncl 0 --batch 1024 --nmd 2048 --steps 1024
$P --label cfg4_shipped_au6 --nd 42 --ncl 15 --ml 100 --batch 4096 --dt 15 --nmd 2048 --burn 128 --steps 512
$P --label cfg4_stretch_nd510_ml2000 --nd 510 --ncl 81 --ml 2000 --batch 1024 --dt 15 --nmd 4096 --burn 2048 --steps 32
$P --label cfg5_nd63_ml1000_B16384 --nd 63 --ncl 15 --ml 1000 --ebath 0 --batch 16384 --nmd 2048 --burn 1024 --steps 256
$P --label cfg5_nd60_ml200_B65536 --nd 60 --ncl 15 --ml 200 --ebath 0 --batch 65536 --nmd 1024 --burn 256 --steps 256
$P --label cfg5_nd129_ml200_B4096 --nd 129 --ncl 30 --ml 200 --ebath 0 --batch 4096 --nmd 512 --burn 256 --steps 64
