"""Quick throughput probe of the stepping path (not the contract bench): AuBZ-like shape."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--nmd", type=int, default=2048)
ap.add_argument("--steps", type=int, default=256)
ap.add_argument("--mode", type=int, default=1)
ap.add_argument("--case", default="aubz")
ap.add_argument("--ml", type=int, default=0)
a = ap.parse_args()
import pymd_testlib as T
c = dict(T.CASES[a.case]); c["nmd"] = a.nmd
if a.ml: c["ml"] = a.ml
t0 = time.time()
m, j = T.device_md(c, a.batch, savepq=False, step_mode=a.mode, seed=1)
m.initialise(); m.ResetHis()
for b in m.baths:
    b.gnoi()
torch.cuda.synchronize()
setup = time.time() - t0
m.steps(32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.steps(a.steps); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
import ctypes
from pymd_b200 import _lib
_lib.call("pymd_profile_enable", m._ctx, 1)
m.steps(a.steps)
ms8 = (ctypes.c_double * 8)(); c8 = (ctypes.c_longlong * 8)()
_lib.call("pymd_profile_read", m._ctx, ctypes.addressof(ms8), ctypes.addressof(c8))
_lib.call("pymd_profile_enable", m._ctx, 0)
names = ["tail_gemm", "fused_step", "other_gemm", "stage", "far_fft"]
prof = {names[i]: dict(us_per_mdstep=round(1e3 * ms8[i] / a.steps, 3), us_per_launch=round(1e3 * ms8[i] / c8[i], 2), n=int(c8[i]))
        for i in range(5) if c8[i]}
print(json.dumps(prof))
print(json.dumps(dict(case=a.case, batch=a.batch, mode=a.mode, steps=a.steps, ms_per_mdstep=ms / a.steps,
                      traj_steps_per_s=a.batch * a.steps / ms * 1e3, setup_s=setup,
                      finite=bool(np.isfinite(m.p).all()))))
