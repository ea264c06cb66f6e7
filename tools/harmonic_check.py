#!/usr/bin/env python3
"""MD power2 of an ensemble in the steady state against the analytic harmonic spectrum
(md.HarmonicSpectra): prints the integrated ratio and the band-resolved ratios."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pymd_testlib as T  # noqa: E402


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "ph2"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    nmd = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    c = dict(T.CASES[case], nmd=nmd)
    m, j = T.device_md(c, B, seed=3)
    m.initialise()
    m.ResetHis()
    acc = 0.0
    for rep in range(3):
        for b in m.baths:
            b.gnoi()
        m.ResetSavepq(True)
        m.steps(nmd)
        if rep > 0:
            m.GetPower()
            acc = acc + m.power2[:, 1] / 2.0
    h = m.HarmonicSpectra()
    k = np.rint(h["w"] / (2 * np.pi / c["dt"] / nmd)).astype(int)
    mdp = acc[k]
    print("case %s  B=%d nmd=%d  nw=%d" % (case, B, nmd, len(k)))
    print("integrated ratio MD/analytic: %.4f" % (mdp.sum() / h["PP"].sum()))
    nb = 8
    edges = np.linspace(0, len(k), nb + 1).astype(int)
    for a, b in zip(edges[:-1], edges[1:]):
        print("  w in [%.4f, %.4f]: ratio %.4f  (PP share %.3f)" % (h["w"][a], h["w"][b - 1], mdp[a:b].sum() / h["PP"][a:b].sum(),
                                                                  h["PP"][a:b].sum() / h["PP"].sum()))


if __name__ == "__main__":
    main()
