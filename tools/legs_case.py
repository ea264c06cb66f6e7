"""Noise generation and power spectra of the cfg2 e2e run alone (for `ncu --metrics gpu__time_duration.sum`):
`python tools/legs_case.py [batch] [nmd]`."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nmd = int(sys.argv[2]) if len(sys.argv) > 2 else 2 ** 15
c = bench.workload("cfg2", 1)
c["nmd"] = nmd
m = bench.build(c, B, 0, True, 7)
m.enable_record_overlay()
for b in m.baths:
    b.spectral_factors()
for rep in range(2):
    for b in m.baths:
        b._gnoi(segment=rep)
    torch.cuda.synchronize()
    m.GetPower()
    torch.cuda.synchronize()
print("ok")
