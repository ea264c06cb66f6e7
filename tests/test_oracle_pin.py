"""Pin the oracle: with the legacy numpy RNG seeded as in make_golden.py the
Python-3 restatement must reproduce the REFERENCE's own outputs (golden *.npz,
minted by running oracle/_ref) bit-for-bit."""
import numpy as np
import pytest

from oracle import pymd_oracle as O
from pymd_b200 import synth


def drive(m, nrep, seed):
    np.random.seed(seed)
    m.initialise()
    m.ResetHis()
    out = {"p0": np.array(m.p), "q0": np.array(m.q)}
    power = np.zeros((m.nmd, 2))
    power2 = np.zeros((m.nmd, 2))
    for j in range(nrep):
        for b in m.baths:
            b.gnoi()
        m.ResetSavepq()
        for _ in range(m.nmd):
            m.vv()
        m.GetPower()
        power = (power * j + m.power) / float(j + 1)
        power2 = (power2 * j + m.power2) / float(j + 1)
        out["ps%d" % j], out["qs%d" % j], out["etot%d" % j] = m.ps.copy(), m.qs.copy(), m.etot.copy()
        for i, b in enumerate(m.baths):
            out["noise%d_%d" % (j, i)] = np.array(b.noise)
            out["cur%d_%d" % (j, i)] = b.cur.copy()
            out["fhis%d_%d" % (j, i)] = m.fhis[i].copy()
    out["power"], out["power2"] = power, power2
    out["p_end"], out["q_end"] = np.array(m.p), np.array(m.q)
    out["phis_end"], out["qhis_end"] = m.phis.copy(), m.qhis.copy()
    return out


def check(out, g, keys=None):
    for k in (keys or out):
        assert k in g, k
        assert out[k].shape == g[k].shape, k
        assert np.array_equal(out[k], g[k]), "%s differs: max abs %g" % (k, np.abs(out[k] - g[k]).max())


def test_ph2_bit_exact(golden):
    c, g = golden.cfg("ph2"), golden("ph2")
    j = synth.synth(12, 3, 3, seed=11)
    assert np.array_equal(j.DynMat, g["dyn_in"]) and np.array_equal(j.SigL, g["SigL"])
    m = O.md(c["dt"], c["nmd"], c["T"], axyz=j.axyz, harmonic=True, dyn=j.DynMat, nrep=c["nrep"])
    pl = O.phbath(c["T"] + c["dT"] / 2, j.cats_l, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=j.SigL, gwl=j.wl)
    pr = O.phbath(c["T"] - c["dT"] / 2, j.cats_r, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=j.SigR, gwl=j.wl)
    pl.gmem(); pr.gmem()
    m.AddBath(pl); m.AddBath(pr)
    out = drive(m, c["nrep"], c["seed"])
    out.update(dyn=m.dyn, hw=m.hw, U=m.U, kernel0=pl.kernel, kernel1=pr.kernel, gamma0=pl.gamma, gamma1=pr.gamma)
    check(out, g)


def test_eph_bit_exact(golden):
    c, g = golden.cfg("eph"), golden("eph")
    j = synth.synth(12, 3, 3, seed=21, constraint=True)
    m = O.md(c["dt"], c["nmd"], c["T"], axyz=j.axyz, harmonic=True, dyn=j.DynMat, nrep=1, constr=j.constr.copy())
    eb = O.ebath(j.cats_e, c["T"], c["dt"], c["nmd"], wmax=c["ewmax"], nw=c["enw"], bias=c["bias"],
                 efric=j.efric, exim=j.exim, exip=j.exip, zeta1=j.zeta1, zeta2=j.zeta2)
    pl = O.phbath(c["T"], j.cats_l, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=j.SigL, gwl=j.wl)
    pr = O.phbath(c["T"], j.cats_r, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=j.SigR, gwl=j.wl)
    pl.gmem(); pr.gmem()
    m.AddBath(eb); m.AddBath(pl); m.AddBath(pr)
    out = drive(m, 1, c["seed"])
    out.update(dyn=m.dyn, kernel1=pl.kernel, kernel2=pr.kernel)
    check(out, g)


def test_debye_T0_bit_exact(golden):
    c, g = golden.cfg("debye"), golden("debye")
    j = synth.synth(9, 3, 3, seed=31)
    m = O.md(c["dt"], c["nmd"], c["T"], axyz=j.axyz, harmonic=True, dyn=j.DynMat, nrep=1)
    pd = O.phbath(c["Tph"], [0, 2], c["debye"], c["nw"], c["dt"], c["nmd"], classical=True, zpmotion=False)
    pd.gmem()
    eb = O.ebath(j.cats_e, c["T"], c["dt"], c["nmd"], wmax=c["ewmax"], nw=c["enw"], bias=c["bias"],
                 efric=j.efric, exim=j.exim, exip=j.exip)
    m.AddBath(pd); m.AddBath(eb)
    out = drive(m, 1, c["seed"])
    out.update(kernel0=pd.kernel)
    check(out, g)


@pytest.mark.filterwarnings("ignore:overflow")
def test_setup_tables_bit_exact(golden):
    c, g = golden.cfg("setup"), golden("setup")
    j = synth.synth(12, 6, 3, seed=41)
    pa = O.phbath(c["T"], j.cats_l, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=j.SigL, gwl=j.wl, eta_ad=c["eta_ad"])
    pa.gmem()
    assert np.array_equal(pa.kernel, g["kernel"])
    assert np.array_equal(pa.gamma, g["gamma_new"]) and np.array_equal(pa.gammaOld, g["gamma_old"])
    fl = np.array([O.flinterp(x, g["fl_xs"], g["fl_ys"]) for x in g["fl_xq"]])
    assert np.array_equal(fl, g["fl_out"])
    # the reference's convention: mirrored about the nearest grid point (SURVEY 8a note)
    assert fl[list(g["fl_xq"]).index(0.9)] == 11.0
    wq = g["wq"]
    assert np.array_equal(np.array([[O.bose(w, T) for w in wq] for T in (0.0, 4.2, 300.0)]), g["bose_tab"])
    eq = np.array([[[O.equ(w, 0.03, T, cl, zp) for w in wq] for T in (0.0, 300.0)]
                   for cl in (False, True) for zp in (True, False)])
    assert np.array_equal(eq, g["equ_tab"])
    assert np.array_equal(O.Fourier1D(g["ft_in"], c["dt"]), g["ft_t2w"])
    assert np.array_equal(O.iFourier1D(g["ft_in"].astype(complex), c["dt"]), g["ft_w2t"])
    assert np.array_equal(O.powerspec(g["traj"], c["dt"], 16), g["powerspec"])
    assert np.array_equal(O.powerspec2(g["traj"], c["dt"], 16), g["powerspec2"])
    Ds = O.Drs(g["h_wl"], g["h_dyn"], g["h_se"])
    assert np.array_equal(O.DDos(g["h_wl"], Ds), g["h_DDos"])
    assert np.array_equal(O.PP(g["h_wl"], Ds, g["h_noise"]), g["h_PP"])
    assert np.array_equal(O.UU(g["h_wl"], Ds, g["h_noise"]), g["h_UU"])


def test_fft_roundtrip_selfcheck():
    """The reference's only numeric self-check (myfft.py:55-76): iFourier1D(Fourier1D(a)) == a."""
    a = np.arange(10).astype(float)
    assert np.abs(O.iFourier1D(O.Fourier1D(a, 0.1), 0.1) - a).max() < 1e-12


def test_injection_equals_sequential_draws(golden):
    """xi injected by [freq, eigen-index] == the reference's sequential draws
    (no draw where an eigenvalue is <= 0)."""
    c = golden.cfg("ph2")
    j = synth.synth(12, 3, 3, seed=11)
    b = O.phbath(c["T"], j.cats_l, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=j.SigL, gwl=j.wl)
    np.random.seed(9)
    b.gnoi()
    n_seq = b.noise.copy()
    # replay: walk the same eigenvalues, place the sequential draws at [f, i]
    np.random.seed(9)
    hl = c["nmd"] // 2
    xi = np.zeros((hl + 1, b.nc))
    for f in range(hl + 1):
        av, _ = np.linalg.eigh(O.phonon_cov(f, b.gamma, b.gwl, b.T, b.wmax, b.dt, b.nmd, False, True))
        for i in range(b.nc):
            if av[i] > 0:
                xi[f, i] = np.random.standard_normal()
    b.gnoi(xi=xi)
    assert np.array_equal(b.noise, n_seq)
