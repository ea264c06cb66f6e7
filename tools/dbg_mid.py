import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import pymd_testlib as T
case, nt = sys.argv[1], sys.argv[2]
os.environ["PYMD_FUSED_NT"] = nt
c = T.CASES[case]; B = 40
out = {}
KNOB = sys.argv[3] if len(sys.argv) > 3 else "PYMD_FAR"
for far in ("1", "0"):
    os.environ[KNOB] = far
    m, j = T.device_md(c, B, step_mode=2)
    xi, r = T.draw(c, m, B, seed=31)
    m.initialise(r=r); m.ResetHis()
    for b, bath in enumerate(m.baths):
        bath.gnoi(xi=xi[0][b])
    m.ResetSavepq()
    m.steps(c["nmd"])
    out[far] = (np.array(m.etot), [np.array(b.cur) for b in m.baths], np.array(m.ps))
e1, c1, p1 = out["1"]; e0, c0, p0 = out["0"]
print("ps", np.abs(p1 - p0).max(), "per traj", np.round(np.abs(p1 - p0).max(axis=(0, 1)) / np.abs(p0).max(), 9))
print("ps per t", np.abs(p1 - p0).max(axis=(1, 2))[:40])
d = np.abs(e1 - e0) / np.abs(e0).max()
print("etot relerr per t (max over traj):", np.round(d.max(axis=1), 3)[:40])
print("etot relerr per traj:", np.round(d.max(axis=0), 3))
for a, b in zip(c1, c0):
    dd = np.abs(a - b) / np.abs(b).max()
    print("cur per t", np.round(dd.max(axis=1), 3)[:40])
