"""Small stepping run for compute-sanitizer: `python tools/san_case.py CASE BATCH MODE [STEPS]`.
Exercises one path (fused / fused2 with PYMD_FUSED_NT=4 / local / per-step pipeline) on a few CTAs."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import torch
import pymd_testlib as T
case, B, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 40
c = dict(T.CASES[case]); c["nmd"] = 64
m, j = T.device_md(c, B, savepq=True, step_mode=mode, seed=1)
m.initialise(); m.ResetHis()
for b in m.baths:
    b.gnoi()
m.ResetSavepq()
m.steps(3)
m.steps(steps - 3)
torch.cuda.synchronize()
assert np.isfinite(m.p).all()
print("san_case ok", case, B, mode, steps)
