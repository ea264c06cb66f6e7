#!/usr/bin/env python3
"""Mint golden vectors by RUNNING THE REFERENCE ITSELF (oracle/_ref).

The reference's own tests hold no golden vectors, so the
oracle is pinned against the reference's outputs on small synthetic junctions:
this script imports the mechanically converted reference modules from
``oracle/_ref`` (built by ``oracle/build_ref.py`` from /root/reference; they
cannot travel, so their outputs are committed here as ``*.npz``), seeds the
legacy numpy global RNG, drives the reference the way its ``tests/test.py``
does (initialise, ResetHis, per segment: gnoi, nmd x vv, GetPower), and stores
inputs + outputs.

Run in the build container only:   python tests/golden/make_golden.py
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ["PYMD_REF_QUIET"] = "1"
warnings.filterwarnings("ignore", category=SyntaxWarning)

from oracle import build_ref  # noqa: E402
from pymd_b200 import synth  # noqa: E402


def _ref():
    if not build_ref.build():
        raise SystemExit("reference sources absent: golden vectors can only be minted in the build container")
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import md as rmd, phbath as rph, ebath as reb, noise as rnoise, spectrum as rspec  # noqa
    import functions as rfun, myfft as rfft, harmonic as rharm  # noqa
    return rmd, rph, reb, rnoise, rspec, rfun, rfft, rharm


def drive(rmd, mdrun, nrep, seed):
    """tests/test.py:138-157 style hand loop; returns per-segment records."""
    np.random.seed(seed)
    mdrun.initialise()
    mdrun.ResetHis()
    out = {"p0": np.array(mdrun.p), "q0": np.array(mdrun.q)}
    power = np.zeros((mdrun.nmd, 2))
    power2 = np.zeros((mdrun.nmd, 2))
    for j in range(nrep):
        for b in mdrun.baths:
            b.gnoi()
        mdrun.ResetSavepq()
        for _ in range(mdrun.nmd):
            mdrun.vv()
        mdrun.GetPower()
        power = (power * j + mdrun.power) / float(j + 1)      # md.py:604-605
        power2 = (power2 * j + mdrun.power2) / float(j + 1)
        out["ps%d" % j] = mdrun.ps.copy()
        out["qs%d" % j] = mdrun.qs.copy()
        out["etot%d" % j] = mdrun.etot.copy()
        for i, b in enumerate(mdrun.baths):
            out["noise%d_%d" % (j, i)] = np.array(b.noise)
            out["cur%d_%d" % (j, i)] = b.cur.copy()
            out["fhis%d_%d" % (j, i)] = mdrun.fhis[i].copy()
    out["power"], out["power2"] = power, power2
    out["p_end"], out["q_end"] = np.array(mdrun.p), np.array(mdrun.q)
    out["phis_end"], out["qhis_end"] = mdrun.phis.copy(), mdrun.qhis.copy()
    return out


def case_ph2(R):
    """Two non-local phonon baths at different temperatures, two segments."""
    rmd, rph = R[0], R[1]
    j = synth.synth(12, 3, 3, seed=11)
    cfg = dict(dt=8.0, nmd=64, T=300.0, dT=20.0, ml=8, nw=40, hwcut=0.03, nrep=2, seed=1234)
    m = rmd.md(cfg["dt"], cfg["nmd"], cfg["T"], axyz=j.axyz, harmonic=True, dyn=j.DynMat, nrep=cfg["nrep"])
    pl = rph.phbath(cfg["T"] + cfg["dT"] / 2, j.cats_l, cfg["hwcut"], cfg["nw"], cfg["dt"], cfg["nmd"],
                    ml=cfg["ml"], sig=j.SigL, gwl=j.wl)
    pr = rph.phbath(cfg["T"] - cfg["dT"] / 2, j.cats_r, cfg["hwcut"], cfg["nw"], cfg["dt"], cfg["nmd"],
                    ml=cfg["ml"], sig=j.SigR, gwl=j.wl)
    pl.gmem(); pr.gmem()
    m.AddBath(pl); m.AddBath(pr)
    out = drive(rmd, m, cfg["nrep"], cfg["seed"])
    out.update(dyn_in=j.DynMat, dyn=m.dyn, hw=m.hw, U=m.U, wl=j.wl, SigL=j.SigL, SigR=j.SigR,
               kernel0=pl.kernel, kernel1=pr.kernel, gamma0=pl.gamma, gamma1=pr.gamma)
    return cfg, out


def case_eph(R):
    """Electron bath at finite bias (all five matrices) + two phonon baths +
    one constraint vector (AuBZ-like), one segment."""
    rmd, rph, reb = R[0], R[1], R[2]
    j = synth.synth(12, 3, 3, seed=21, constraint=True)
    cfg = dict(dt=1.0, nmd=32, T=100.0, dT=0.0, ml=6, nw=30, hwcut=0.03, nrep=1, seed=77,
               bias=0.3, ewmax=1.0, enw=500)
    m = rmd.md(cfg["dt"], cfg["nmd"], cfg["T"], axyz=j.axyz, harmonic=True, dyn=j.DynMat,
               nrep=1, constr=j.constr.copy())
    eb = reb.ebath(j.cats_e, cfg["T"], cfg["dt"], cfg["nmd"], wmax=cfg["ewmax"], nw=cfg["enw"],
                   bias=cfg["bias"], efric=j.efric, exim=j.exim, exip=j.exip, zeta1=j.zeta1, zeta2=j.zeta2)
    pl = rph.phbath(cfg["T"], j.cats_l, cfg["hwcut"], cfg["nw"], cfg["dt"], cfg["nmd"], ml=cfg["ml"],
                    sig=j.SigL, gwl=j.wl)
    pr = rph.phbath(cfg["T"], j.cats_r, cfg["hwcut"], cfg["nw"], cfg["dt"], cfg["nmd"], ml=cfg["ml"],
                    sig=j.SigR, gwl=j.wl)
    pl.gmem(); pr.gmem()
    m.AddBath(eb); m.AddBath(pl); m.AddBath(pr)
    out = drive(rmd, m, 1, cfg["seed"])
    out.update(dyn_in=j.DynMat, dyn=m.dyn, wl=j.wl, SigL=j.SigL, SigR=j.SigR, constr=j.constr,
               efric=j.efric, exim=j.exim, exip=j.exip, zeta1=j.zeta1, zeta2=j.zeta2,
               kernel1=pl.kernel, kernel2=pr.kernel)
    return cfg, out


def case_debye(R):
    """Time-local Debye bath (phbath.py:82-89), classical noise without
    zero-point motion, plus a T = 0 electron bath with bias (Alkane-like)."""
    rmd, rph, reb = R[0], R[1], R[2]
    j = synth.synth(9, 3, 3, seed=31)
    cfg = dict(dt=2.0, nmd=32, T=0.0, Tph=50.0, debye=0.02, nw=20, nrep=1, seed=5, bias=0.5,
               ewmax=1.0, enw=100)
    m = rmd.md(cfg["dt"], cfg["nmd"], cfg["T"], axyz=j.axyz, harmonic=True, dyn=j.DynMat, nrep=1)
    pd = rph.phbath(cfg["Tph"], [0, 2], cfg["debye"], cfg["nw"], cfg["dt"], cfg["nmd"],
                    classical=True, zpmotion=False)
    pd.gmem()
    eb = reb.ebath(j.cats_e, cfg["T"], cfg["dt"], cfg["nmd"], wmax=cfg["ewmax"], nw=cfg["enw"],
                   bias=cfg["bias"], efric=j.efric, exim=j.exim, exip=j.exip)
    m.AddBath(pd); m.AddBath(eb)
    out = drive(rmd, m, 1, cfg["seed"])
    out.update(dyn_in=j.DynMat, efric=j.efric, exim=j.exim, exip=j.exip, kernel0=pd.kernel)
    return cfg, out


def case_setup(R):
    """Setup-time tables only: artificial-damping kernel (eta_ad != 0), the
    flinterp mirror convention, bose/equ special cases, FT conventions,
    spectra, analytic harmonic Green's functions."""
    rph, rnoise, rspec, rfun, rfft, rharm = R[1], R[3], R[4], R[5], R[6], R[7]
    j = synth.synth(12, 6, 3, seed=41)
    cfg = dict(dt=4.0, nmd=16, T=150.0, ml=5, nw=25, hwcut=0.03, eta_ad=1.0e-3)
    pa = rph.phbath(cfg["T"], j.cats_l, cfg["hwcut"], cfg["nw"], cfg["dt"], cfg["nmd"], ml=cfg["ml"],
                    sig=j.SigL, gwl=j.wl, eta_ad=cfg["eta_ad"])
    pa.gmem()
    xs = np.array([0.0, 1.0, 2.0, 3.0])
    ys = np.array([0.0, 10.0, 20.0, 40.0])
    xq = np.array([-0.5, 0.0, 0.2, 0.49, 0.5, 0.51, 0.9, 1.0, 1.3, 1.5, 2.2, 2.6, 3.0, 3.7])
    fl = np.array([rfun.flinterp(x, xs, ys) for x in xq])
    wq = np.array([-0.3, -1e-3, 0.0, 1e-4, 0.01, 0.029, 0.03, 0.05])
    bose_tab = np.array([[rfun.bose(w, T) for w in wq] for T in (0.0, 4.2, 300.0)])
    equ_tab = np.array([[[rnoise.equ(w, 0.03, T, cl, zp) for w in wq] for T in (0.0, 300.0)]
                        for cl in (False, True) for zp in (True, False)])
    rng = np.random.default_rng(3)
    a = rng.standard_normal(16)
    f = rfft.myfft(cfg["dt"], 16)
    traj = rng.standard_normal((16, 6))
    wl = np.linspace(0.001, 0.03, 7)
    se = np.zeros((7, 12, 12), complex)
    se[:, :6, :6] = np.array([rfun.flinterp(w, j.wl, j.SigL) for w in wl])
    Ds = rharm.Drs(wl, j.DynMat, se)
    pn = np.zeros((7, 12, 12), complex)
    gam = np.array([-rfun.flinterp(w, j.wl, j.SigL).imag / w for w in wl])
    pn[:, :6, :6] = rnoise.phnoisew(gam, wl, cfg["T"], cfg["hwcut"])
    out = dict(wl=j.wl, SigL=j.SigL, kernel=pa.kernel, gamma_new=pa.gamma, gamma_old=pa.gammaOld,
               fl_xs=xs, fl_ys=ys, fl_xq=xq, fl_out=fl, wq=wq, bose_tab=bose_tab, equ_tab=equ_tab,
               ft_in=a, ft_t2w=f.Fourier1D(a), ft_w2t=f.iFourier1D(a.astype(complex)),
               traj=traj, powerspec=rspec.powerspec(traj, cfg["dt"], 16),
               powerspec2=rspec.powerspec2(traj, cfg["dt"], 16),
               h_wl=wl, h_dyn=j.DynMat, h_se=se, h_noise=pn, h_DDos=rharm.DDos(wl, Ds),
               h_PP=rharm.PP(wl, Ds, pn), h_UU=rharm.UU(wl, Ds, pn))
    return cfg, out


def main():
    R = _ref()
    manifest = {"numpy": np.__version__, "generator": "tests/golden/make_golden.py",
                "source": "oracle/_ref (mechanical py2->py3 conversion of /root/reference)",
                "cases": {}}
    for name, fn in (("ph2", case_ph2), ("eph", case_eph), ("debye", case_debye), ("setup", case_setup)):
        cfg, out = fn(R)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        manifest["cases"][name] = cfg
        print(name, "->", sorted(out)[:6], "...")
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
