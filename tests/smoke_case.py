"""__graft_entry__.smoke(): one small ensemble run of the hot path on cuda:0 -- noise
generation (tensor-pipe shaping + FFT), 64 velocity-Verlet steps with an electron bath, two
non-local phonon baths with a 40-step memory (fused step kernel with in-kernel mid field + FFT far
field) and a constraint, power spectra -- checked against the oracle (test infrastructure)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def run(verbose=True):
    import torch
    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    import pymd_testlib as T
    c = dict(T.CASES["eph"], ml=40, nmd=64)
    B = 5
    m, j = T.device_md(c, B)
    xi, r = T.draw(c, m, B, seed=123)
    m.initialise(r=r)
    m.ResetHis()
    for b, bath in enumerate(m.baths):
        bath.gnoi(xi=xi[0][b])
    m.ResetSavepq()
    m.steps(c["nmd"])
    m.GetPower()
    torch.cuda.synchronize()
    worst = 0.0
    pw = np.zeros(c["nmd"])
    for n in range(B):
        om, recs = T.run_oracle(c, j, m, xi, r, n, c["nmd"])
        rec = recs[0]
        worst = max(worst, T.relerr(m.ps[..., n], rec["ps"]), T.relerr(m.qs[..., n], rec["qs"]),
                    T.relerr(m.etot[..., n], rec["etot"]))
        for b in range(len(m.baths)):
            worst = max(worst, T.relerr(m.baths[b].cur[:, n], rec["cur"][b]))
        om.GetPower()
        pw += om.power2[:, 1]
    worst = max(worst, T.relerr(m.power2[:, 1], pw / B))
    if verbose:
        print("smoke: ensemble of %d, %d steps, max rel. deviation from the oracle %.3e" % (B, c["nmd"], worst))
    assert worst < 1e-10, worst
    return worst


if __name__ == "__main__":
    run()
