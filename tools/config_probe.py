"""Throughput probe of the other BASELINE configs (mode auto): cfg3 Alkane-like (ebath only, nd=96,
bias), cfg4 Au6-like stretch (nd=510, leads nc=81, ml=2000), and a sweep point; this only reports traj-steps/s
(bench.py --config ... is the driver-visible way).  Parity of these shapes: tests/test_gpu_trajectory.py
test_baseline_config_shapes (nd = 96, nd = 42), test_large_junctions_with_memory_baths (nd = 129 / 150 with memory
baths), test_delay_line_far_field_of_the_per_step_pipeline."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from pymd_b200 import synth
from pymd_b200.ebath import ebath
from pymd_b200.md import md
from pymd_b200.phbath import phbath

ap = argparse.ArgumentParser()
ap.add_argument("--nd", type=int, default=96)
ap.add_argument("--ncl", type=int, default=0)
ap.add_argument("--ml", type=int, default=100)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--nmd", type=int, default=256)
ap.add_argument("--steps", type=int, default=64)
ap.add_argument("--ebath", type=int, default=1)
ap.add_argument("--dt", type=float, default=2.0)
ap.add_argument("--burn", type=int, default=8, help="untimed steps before the timed window (>= ml fills the history)")
ap.add_argument("--nce", type=int, default=0, help="electron-bath dofs (default: all)")
ap.add_argument("--label", default="")
a = ap.parse_args()
ncl = a.ncl
j = synth.synth(a.nd, max(ncl, 3), max(ncl, 3), seed=9, nc_e=(a.nce or a.nd) if a.ebath else None)
t0 = time.time()
m = md(a.dt, a.nmd, 300.0, axyz=j.axyz, harmonic=True, dyn=j.DynMat, savepq=False, batch=a.batch, seed=3)
m.write_files = False
if a.ebath:
    m.AddBath(ebath(j.cats_e, 300.0, a.dt, a.nmd, wmax=1.0, nw=500, bias=0.5, efric=j.efric, exim=j.exim, exip=j.exip,
                    zeta1=j.zeta1, zeta2=j.zeta2))
if ncl:
    for T, cats, sig in ((305.0, j.cats_l, j.SigL), (295.0, j.cats_r, j.SigR)):
        b = phbath(T, cats, 0.03, 200, a.dt, a.nmd, ml=a.ml, sig=sig, gwl=j.wl)
        b.gmem(); m.AddBath(b)
m.initialise(); m.ResetHis()
for b in m.baths:
    b.gnoi()
torch.cuda.synchronize()
setup = time.time() - t0
m.steps(a.burn)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.steps(a.steps); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps(dict(label=a.label, nd=a.nd, ncl=ncl, ml=a.ml, ebath=a.ebath, batch=a.batch, ms_per_mdstep=ms / a.steps,
                      traj_steps_per_s=a.batch * a.steps / ms * 1e3, setup_s=round(setup, 1),
                      finite=bool(np.isfinite(m.p).all()))))
