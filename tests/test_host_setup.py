"""Host-side setup tables of pymd_b200 (vectorised numpy) against the golden vectors minted
from the reference and against the oracle.  Tolerances: these tables differ from the
reference only in summation order (vectorised mean / tensordot), so 1e-13 relative to the
largest entry; element-wise functions must agree exactly."""
import numpy as np
import pytest

from oracle import pymd_oracle as O
from pymd_b200 import synth
from pymd_b200 import functions as F
from pymd_b200 import noise as PN
from pymd_b200 import harmonic as H
from pymd_b200.phbath import phbath, gamt
from pymd_b200.ebath import ebath
from pymd_b200.md import md, ApplyConstraint, sameq


def relmax(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.mark.filterwarnings("ignore:overflow")
def test_scalar_tables_exact(golden):
    g = golden("setup")
    assert np.array_equal(F.flinterp(g["fl_xq"], g["fl_xs"], g["fl_ys"]), g["fl_out"])
    for x, y in zip(g["fl_xq"], g["fl_out"]):
        assert F.flinterp(float(x), g["fl_xs"], g["fl_ys"]) == y
    wq = g["wq"]
    assert np.array_equal(np.array([F.bose(wq, T) for T in (0.0, 4.2, 300.0)]), g["bose_tab"])
    eq = np.array([[PN.equ(wq, 0.03, T, cl, zp) for T in (0.0, 300.0)]
                   for cl in (False, True) for zp in (True, False)])
    assert np.array_equal(eq, g["equ_tab"])
    assert F.bose(0.01, 300.0) == O.bose(0.01, 300.0) and F.bose(0.0, 300.0) == 0.0
    assert F.nearest(0.9, [0, 1, 2]) == 1


def test_flinterp_matrix_valued_matches_oracle():
    j = synth.synth(12, 6, 3, seed=41)
    gam = -np.imag(j.SigL[1:]) / j.wl[1:, None, None]
    xs = j.wl[1:]
    xq = np.concatenate(([0.0, xs[0], xs[-1], 1.0], np.random.default_rng(0).uniform(0, 0.07, 50)))
    got = F.flinterp(xq, xs, gam)
    for i, x in enumerate(xq):
        assert np.array_equal(got[i], O.flinterp(x, xs, gam))


def test_kernel_tables_vs_reference(golden):
    c, g = golden.cfg("ph2"), golden("ph2")
    j = synth.synth(12, 3, 3, seed=11)
    pl = phbath(c["T"] + c["dT"] / 2, j.cats_l, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=j.SigL, gwl=j.wl)
    pl.gmem()
    assert np.array_equal(pl.gamma, g["gamma0"])
    assert pl.kernel.shape == g["kernel0"].shape and relmax(pl.kernel, g["kernel0"]) < 1e-13
    assert list(pl.cids) == [0, 1, 2] and pl.nc == 3 and not pl.local and pl.wmax == 0.06


def test_eta_ad_kernel_vs_reference(golden):
    c, g = golden.cfg("setup"), golden("setup")
    j = synth.synth(12, 6, 3, seed=41)
    pa = phbath(c["T"], j.cats_l, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=j.SigL, gwl=j.wl,
                eta_ad=c["eta_ad"])
    pa.gmem()
    assert relmax(pa.kernel, g["kernel"]) < 1e-13
    assert relmax(pa.gamma, g["gamma_new"]) < 1e-13
    assert np.array_equal(pa.gammaOld, g["gamma_old"])


def test_debye_bath_is_local(golden):
    c, g = golden.cfg("debye"), golden("debye")
    pd = phbath(c["Tph"], [0, 2], c["debye"], c["nw"], c["dt"], c["nmd"], classical=True, zpmotion=False)
    pd.gmem()
    assert pd.local and pd.ml == 1 and np.array_equal(pd.kernel, g["kernel0"])


def test_ebath_matrices_and_errors(golden):
    g = golden("eph")
    c = golden.cfg("eph")
    j = synth.synth(12, 3, 3, seed=21, constraint=True)
    eb = ebath(j.cats_e, c["T"], c["dt"], c["nmd"], wmax=1.0, nw=500, bias=c["bias"], efric=j.efric, exim=j.exim,
               exip=j.exip, zeta1=j.zeta1, zeta2=j.zeta2)
    ob = O.ebath(j.cats_e, c["T"], c["dt"], c["nmd"], wmax=1.0, nw=500, bias=c["bias"], efric=j.efric, exim=j.exim,
                 exip=j.exip, zeta1=j.zeta1, zeta2=j.zeta2)
    for k in ("efric", "exim", "exip", "zeta1", "zeta2"):
        assert np.array_equal(getattr(eb, k), getattr(ob, k))
    assert eb.ml == 1 and eb.nc == 12 and eb.kernel.shape == (1, 12, 12)
    eb.GetSig(); ob.GetSig()
    assert np.allclose(eb.sig, ob.sig, rtol=1e-15, atol=0)
    with pytest.raises(ValueError):
        ebath(j.cats_e, 1.0, 1.0, 16, efric=np.eye(5))
    assert ebath([0], 1.0, 1.0, 16).ebath is False


def test_spectral_factors_reproduce_covariance(golden):
    """L L^dagger == covariance wherever it is PSD, and lam/U equal numpy.linalg.eigh of the
    oracle's covariance matrices (same LAPACK call => same gauge)."""
    c = golden.cfg("eph")
    j = synth.synth(12, 3, 3, seed=21)
    pl = phbath(c["T"], j.cats_l, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=j.SigL, gwl=j.wl)
    fac = pl.spectral_factors()
    for f in range(c["nmd"] // 2 + 1):
        cov = O.phonon_cov(f, pl.gamma, pl.gwl, pl.T, pl.wmax, pl.dt, pl.nmd, False, True)
        av, au = np.linalg.eigh(cov)
        assert np.array_equal(fac.lam[f], av) and np.array_equal(fac.U[f], au)
        assert np.allclose(fac.L[f] @ fac.L[f].conj().T, (au * np.maximum(av, 0)) @ au.T, rtol=1e-12, atol=1e-30)
    eb = ebath(j.cats_e, c["T"], c["dt"], c["nmd"], wmax=1.0, nw=500, bias=c["bias"], efric=j.efric, exim=j.exim,
               exip=j.exip)
    fe = eb.spectral_factors()
    assert fe.complex
    for f in (0, 1, 7, c["nmd"] // 2):
        cov = O.electron_cov(f, eb.efric, eb.exim, eb.exip, eb.bias, eb.T, eb.wmax, eb.dt, eb.nmd, False, True)
        assert relmax(fe.lam[f], np.linalg.eigvalsh(cov)) < 1e-12
        pos = (fe.U[f] * np.maximum(fe.lam[f], 0)) @ fe.U[f].conj().T
        assert np.allclose(fe.L[f] @ fe.L[f].conj().T, pos, rtol=1e-12, atol=1e-30)


def test_noise_spectra_and_harmonic_vs_reference(golden):
    g = golden("setup")
    Ds = H.Drs(g["h_wl"], g["h_dyn"], g["h_se"])
    assert relmax(H.DDos(g["h_wl"], Ds), g["h_DDos"]) < 1e-12
    assert relmax(H.PP(g["h_wl"], Ds, g["h_noise"]), g["h_PP"]) < 1e-12
    assert relmax(H.UU(g["h_wl"], Ds, g["h_noise"]), g["h_UU"]) < 1e-12
    j = synth.synth(12, 6, 3, seed=41)
    wl = g["h_wl"]
    gam = np.array([-O.flinterp(w, j.wl, j.SigL).imag / w for w in wl])
    assert np.array_equal(PN.phnoisew(gam, wl, 150.0, 0.03), O.phnoisew(gam, wl, 150.0, 0.03))
    a = PN.enoisew(wl, j.efric, j.exim, j.exip, 0.3, 100.0, 1.0)
    b = O.enoisew(wl, j.efric, j.exim, j.exip, 0.3, 100.0, 1.0)
    assert relmax(a, b) < 1e-14
    assert relmax(H.SigE(wl, j.efric, j.exim, j.zeta1, j.zeta2, 0.3), O.SigE(wl, j.efric, j.exim, j.zeta1, j.zeta2, 0.3)) < 1e-15


def test_setdyn_and_helpers(golden):
    g = golden("ph2")
    m = md(8.0, 64, 300.0, dyn=g["dyn_in"])
    assert relmax(m.dyn, g["dyn"]) < 1e-13 and relmax(m.hw, g["hw"]) < 1e-12
    assert m.nph == 12 and m.ml == 1 and m.t == 0
    bad = g["dyn_in"].copy()
    bad[0, 0] = -1.0                       # negative eigenvalue must be clamped (md.py:228-237)
    m2 = md(8.0, 64, 300.0, dyn=bad)
    assert m2.hw.min() >= 0 and np.linalg.eigvalsh(m2.dyn).min() > -1e-18
    c = np.array([[1.0, 2.0, 0.0], [0.0, 0.0, 3.0]])
    f = np.array([0.3, -0.2, 0.9])
    assert np.array_equal(ApplyConstraint(f, c), O.ApplyConstraint(f, c))
    assert ApplyConstraint(f, None) is f
    assert sameq([1.0, 2.0], [1.0, 2.0 + 1e-10]) and not sameq([1.0], [1.0, 2.0]) and not sameq([1.0], [1.1])
    with pytest.raises(ValueError):
        md(8.0, 64, 300.0, axyz=[["Au", 0, 0, 0]], dyn=g["dyn_in"])
    with pytest.raises(ValueError):
        m.AddBath(phbath(300.0, [0], 0.03, 10, 4.0, 64))        # dt mismatch (md.py:164-166)


def test_gamt_vs_oracle_large_table():
    j = synth.synth(12, 6, 3, seed=5)
    gam = -np.imag(j.SigL[1:]) / j.wl[1:, None, None]
    wl = [0.06 * i / 60 for i in range(60)]
    tl = [4.0 * i for i in range(12)]
    assert relmax(gamt(tl, wl, j.wl[1:], gam), O.gamt(tl, wl, j.wl[1:], gam)) < 1e-13


def test_flinterp_weights_reproduce_flinterp():
    """flinterp as two table rows and weights (what pymd_eigh_factors assembles in the kernel)."""
    from pymd_b200.functions import flinterp, flinterp_weights
    rng = np.random.default_rng(0)
    xs = np.sort(rng.uniform(0, 1, 20))
    ys = rng.standard_normal((20, 3, 3))
    x = np.concatenate([xs, rng.uniform(-0.2, 1.2, 400), 0.5 * (xs[1:] + xs[:-1])])
    idx, wt = flinterp_weights(x, xs)
    assert idx.dtype == np.int32 and idx.shape == wt.shape == (len(x), 2)
    got = wt[:, 0, None, None] * ys[idx[:, 0]] + wt[:, 1, None, None] * ys[idx[:, 1]]
    ref = np.array([O.flinterp(v, xs, ys) for v in x])
    assert relmax(got, ref) < 1e-14 and relmax(flinterp(x, xs, ys), ref) < 1e-15
    i1, w1 = flinterp_weights([0.3], [0.0])                # one-point table (Debye bath)
    assert i1.tolist() == [[0, 0]] and w1[0, 0] == 1.0 and w1[0, 1] == 0.0


def test_bath_selfenergy_and_noise_spectrum_helpers():
    """What md.HarmonicSpectra feeds the harmonic kernel: the discrete-kernel self-energy of a lead,
    the wide-band one of the electron bath (harmonic.SigE), and the noise spectra."""
    j = synth.synth(12, 3, 3, seed=1)
    b = phbath(300.0, j.cats_l, 0.03, 40, 8.0, 64, ml=8, sig=j.SigL, gwl=j.wl)
    b.gmem()
    w = np.linspace(0.001, 0.02, 5)
    se = b.effective_selfenergy(w)
    for i, wv in enumerate(w):
        ref = sum(-1j * wv * b.dt * b.kernel[k] * np.exp(1j * wv * k * b.dt) for k in range(b.ml))
        assert relmax(se[i], ref) < 1e-14
    gam = np.array([O.flinterp(v, b.gwl, b.gamma) for v in w])
    assert relmax(b.noise_spectrum(w), O.phnoisew(gam, w, 300.0, b.wmax)) < 1e-15
    e = ebath(j.cats_e, 100.0, 1.0, 64, wmax=1.0, nw=500, bias=0.3, efric=j.efric, exim=j.exim, exip=j.exip,
              zeta1=j.zeta1, zeta2=j.zeta2)
    assert relmax(e.effective_selfenergy(w), O.SigE(w, j.efric, j.exim, j.zeta1, j.zeta2, 0.3)) < 1e-15
    assert relmax(e.noise_spectrum(w), O.enoisew(w, j.efric, j.exim, j.exip, 0.3, 100.0, 1.0)) < 1e-14
    d = phbath(300.0, [0], 0.03, 40, 8.0, 64)
    d.gmem()
    assert relmax(d.effective_selfenergy(w), -1j * w[:, None, None] * d.kernel[0][None]) < 1e-15


def test_langevin_current_formula_reduces_to_landauer():
    """examples/thermal_conductance.py: the stationary current of the linear Langevin equation,
    J_b = 2 int dw/2pi Re Tr[i w D (S_b + S D^+ Sigma_b^+)], conserves energy (J_L + J_R = 0) and equals
    Landauer's formula when the noise spectra are the fluctuation-dissipation partners
    S_b = (n_b + 1/2) Gamma_b of the friction kernels; with the spectra the baths really generate it
    differs by a few per cent (the number the GPU test compares the MD ensemble with)."""
    import importlib.util
    import os
    import types
    from pymd_b200.functions import bose
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("thermal_conductance_cpu", os.path.join(root, "examples", "thermal_conductance.py"))
    tc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tc)
    j = synth.synth(12, 3, 3, seed=4)
    phl = phbath(160.0, j.cats_l, 0.03, 200, 8.0, 1024, ml=100, sig=j.SigL, gwl=j.wl)
    phr = phbath(140.0, j.cats_r, 0.03, 200, 8.0, 1024, ml=100, sig=j.SigR, gwl=j.wl)
    for b in (phl, phr):
        b.gmem()
    m = types.SimpleNamespace(baths=[phl, phr], nph=12, dyn=j.DynMat)     # what the two functions read from md
    jland, w, trans = tc.landauer_current(m, 160.0, 140.0, nw=1500)
    assert jland > 0 and trans.max() < 3.0 + 1e-6 and trans.min() > -1e-9   # three channels at most
    jl, jr = tc.langevin_current(m, nw=1500)
    assert abs(jl + jr) < 1e-12 * abs(jl)
    assert 0.0 < abs(0.5 * (jl - jr) / jland - 1.0) < 0.15                  # not an exact FDT pair
    for b in (phl, phr):                                                    # make it one
        def fdt(wv, b=b):
            s = b.effective_selfenergy(wv)
            gam = 1j * (s - np.conj(np.transpose(s, (0, 2, 1))))
            return (bose(wv, b.T) + 0.5)[:, None, None] * gam
        b.noise_spectrum = fdt
    jl, jr = tc.langevin_current(m, nw=1500)
    assert abs(0.5 * (jl - jr) / jland - 1.0) < 1e-10 and abs(jl + jr) < 1e-12 * abs(jl)
