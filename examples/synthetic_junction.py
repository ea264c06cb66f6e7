#!/usr/bin/env python3
"""A driver script in the style of the reference's examples (examples/AuC2.py, examples/AuBZ.py),
run on synthetic input files because the SIESTA / Inelastica outputs those scripts read do not ship.

  python examples/synthetic_junction.py [--batch 256] [--nmd 1024] [--nrep 1] [--outdir run]

Step 1 writes the three NetCDF files an AuC2-style run reads (PHrun's Dev.nc, avLambda.nc, Sig.nc);
step 2 is the reference script body with only the import lines changed: read the files, build
EPH.nc, attach an electron bath at finite bias and two phonon leads, and call mdrun.Run().
"""
import argparse
import os
import sys

import numpy as N

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

# -- the only lines that differ from a reference script:  from md import * ; from phbath import * ; ...
from pymd_b200.md import *          # noqa: F401,F403
from pymd_b200.phbath import *      # noqa: F401,F403
from pymd_b200.ebath import *       # noqa: F401,F403
from pymd_b200.myio import *        # noqa: F401,F403
from pymd_b200.matrix import mm
from pymd_b200 import synth


def write_inputs(outdir, nd=48, ncl=15, ncr=15, seed=1):
    """Synthetic stand-ins for ../PHrun/Dev*.nc, ../Lambda*/avLambda.nc and Sig.nc."""
    from scipy.io import netcdf_file
    j = synth.synth(nd, ncl, ncr, seed=seed, nc_e=nd)
    hw2, vec = N.linalg.eigh(j.DynMat)
    hw, U = N.sqrt(N.abs(hw2)), vec.T                    # rows of U are modes: dyn = U^T diag(hw^2) U
    f = netcdf_file(os.path.join(outdir, "Dev.nc"), "w")
    f.createDimension("n", nd)
    v = f.createVariable("hw", "d", ("n",)); v[:] = hw
    v = f.createVariable("U", "d", ("n", "n")); v[:] = U
    f.close()
    f = netcdf_file(os.path.join(outdir, "avLambda.nc"), "w")
    f.createDimension("two", 2); f.createDimension("n", nd)
    v = f.createVariable("muLR", "d", ("two",)); v[:] = [0.25, -0.25]
    for name, a in (("eta", j.efric), ("xim", j.exim), ("xip", j.exip), ("zeta1", j.zeta1), ("zeta2", j.zeta2)):
        v = f.createVariable(name, "d", ("n", "n")); v[:] = U @ a @ U.T          # stored in the mode basis
    f.close()
    WriteEPHNCfile(os.path.join(outdir, "Sig.nc"), j.wl, hw, U, j.DynMat, j.SigL, j.SigR, 0 * j.efric, 0 * j.efric,
                   0 * j.efric, 0 * j.efric, 0 * j.efric)
    return j


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--nmd", type=int, default=1024)
    ap.add_argument("--nrep", type=int, default=1)
    ap.add_argument("--outdir", default="run_synthetic")
    ap.add_argument("--harmonic-check", action="store_true",
                    help="also write harmonic.<T>.run<j>.dat: analytic PP next to the MD power2")
    a = ap.parse_args()
    os.makedirs(a.outdir, exist_ok=True)
    j = write_inputs(a.outdir)

    # ---------------------------------------------------------------- reference-style script body
    T, dT, dt, nmd, hwcut = 300.0, 20.0, 1.0, a.nmd, 0.03
    dyn, U, hw = ReadDynmat(os.path.join(a.outdir, "Dev.nc"))
    eV, eta, xim, xip, zeta1, zeta2 = ReadwbLambda(os.path.join(a.outdir, "avLambda.nc"))
    eph = ReadSig(os.path.join(a.outdir, "Sig.nc"))
    # to real space (examples/AuC2.py:135-139)
    eta, xim, xip, zeta1, zeta2 = (mm(U.T, x, U) for x in (eta, xim, xip, zeta1, zeta2))
    WriteEPHNCfile(os.path.join(a.outdir, "EPH.nc"), eph.wl, hw, U, dyn, eph.SigL, eph.SigR, eta, xim, xip, zeta1, zeta2)

    mdrun = md(dt, nmd, T, axyz=j.axyz, harmonic=True, dyn=dyn, nrep=a.nrep, npie=4,
               batch=a.batch, seed=2024)                               # batch/seed: the only new arguments
    mdrun.outdir = a.outdir
    mdrun.harmonic_check = a.harmonic_check
    ecats = range(mdrun.na)
    eb = ebath(ecats, T, mdrun.dt, mdrun.nmd, wmax=1.0, nw=500, bias=eV, efric=eta, exim=xim, exip=xip,
               zeta1=zeta1, zeta2=zeta2)
    mdrun.AddBath(eb)
    phl = phbath(T + dT / 2, j.cats_l, hwcut, 200, mdrun.dt, mdrun.nmd, ml=100, sig=eph.SigL, gwl=eph.wl)
    phl.gmem()
    mdrun.AddBath(phl)
    phr = phbath(T - dT / 2, j.cats_r, hwcut, 200, mdrun.dt, mdrun.nmd, ml=100, sig=eph.SigR, gwl=eph.wl)
    phr.gmem()
    mdrun.AddBath(phr)
    mdrun.Run()
    # ----------------------------------------------------------------------------------------------
    cur = N.asarray(mdrun.curmean)[-1]
    print("mean currents (nW): electrons %.4g, left lead %.4g, right lead %.4g" % tuple(cur))
    print("outputs in %s: %s" % (a.outdir, ", ".join(sorted(os.listdir(a.outdir)))))
    return mdrun


if __name__ == "__main__":
    main()
