import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import pymd_testlib as T
case, ml, nmd = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
c = dict(T.CASES[case]); c["nmd"] = nmd; c["ml"] = ml
for kv in sys.argv[4:]:
    k_, v_ = kv.split("=")
    c[k_] = (v_ == "True") if v_ in ("True", "False") else (None if v_ == "None" else int(v_))
B = 9
out = {}
for key, env in (("f2_far", dict(PYMD_FUSED_NT="4", PYMD_FUSED2="1", PYMD_FAR="1")), ("f2_nofar", dict(PYMD_FUSED_NT="4", PYMD_FUSED2="1", PYMD_FAR="0")),
                 ("f1_far", dict(PYMD_FUSED_NT="4", PYMD_FUSED2="0", PYMD_FAR="1")), ("nt1_far", dict(PYMD_FUSED_NT="1", PYMD_FUSED2="0", PYMD_FAR="1"))):
    os.environ.update(env)
    m, j = T.device_md(c, B, step_mode=2)
    xi, r = T.draw(c, m, B, seed=13)
    m.initialise(r=r); m.ResetHis()
    for b, bath in enumerate(m.baths):
        bath.gnoi(xi=xi[0][b])
    m.ResetSavepq()
    done = 0
    for n in [3, 8, 29, 1, 64, 7, 100, 40]:
        n = min(n, nmd - done)
        if n:
            m.steps(n); done += n
    m.steps(nmd - done)
    out[key] = np.array(m.ps)
ref = out["nt1_far"]
for k, v in out.items():
    d = np.abs(v - ref).max(axis=(1, 2)) / np.abs(ref).max()
    first = int(np.argmax(d > 1e-10)) if (d > 1e-10).any() else -1
    print(k, "max rel diff vs nt1_far %.2e" % d.max(), "first bad t", first)
