"""Short stepping run for ncu: AuBZ shape, B trajectories, a few fused blocks."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
import pymd_testlib as T
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = dict(T.CASES["aubz"]); c["nmd"] = 256
m, j = T.device_md(c, B, savepq=False, step_mode=mode, seed=1)
m.initialise(); m.ResetHis()
for b in m.baths:
    b.gnoi()
m.steps(128)          # history full (ml = 100)
torch.cuda.synchronize()
m.steps(24)
torch.cuda.synchronize()
print("ok")
