"""Set-up tables on the device (csrc/setup.cu) against the oracle: the friction kernel in time
(phbath.gamt, phbath.py:221-254, with and without artificial damping) and the analytic harmonic
spectra (harmonic.py:37-84).  Both are floating point with a different summation order than
numpy, so the tolerances are relative and stated per test."""
import numpy as np
import pytest

from oracle import pymd_oracle as O
from pymd_b200 import harmonic as H
from pymd_b200 import synth
from pymd_b200.phbath import cosine_transform_device, gamt_device, phbath

pytestmark = pytest.mark.gpu


def lead_gamma(j):
    return -np.imag(j.SigL[1:]) / j.wl[1:, None, None], j.wl[1:]


@pytest.mark.parametrize("ml,nw,nc_l,eta", [(12, 60, 6, 0.0), (37, 201, 15, 0.0), (100, 200, 15, 0.0),
                                             (33, 57, 3, 2e-3), (257, 203, 9, 5e-4), (2000, 200, 27, 0.0)])
def test_gamt_device_vs_oracle(cuda, ml, nw, nc_l, eta):
    """|K_dev - K_oracle| <= 1e-12 max|K|: ml, nw, nc^2 that are not multiples of the tile sizes."""
    j = synth.synth(max(2 * nc_l, 12), nc_l, 3, seed=ml)
    gam, gwl = lead_gamma(j)
    wl = [0.06 * i / nw for i in range(nw)]
    tl = [4.0 * i for i in range(ml)]
    ref = np.real(O.gamt(tl, wl, gwl, gam, eta))
    got = gamt_device(tl, wl, gwl, gam, eta, cuda)
    assert got.shape == ref.shape == (ml, nc_l, nc_l)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()


def test_gmem_device_matches_host_tables(cuda):
    """phbath.gmem(device=True): the kernel and, with eta_ad, the re-derived gamma follow the host path."""
    j = synth.synth(24, 9, 6, seed=3)
    for eta in (0.0, 1e-3):
        host = phbath(300.0, j.cats_l, 0.03, 200, 4.0, 256, ml=120, sig=j.SigL, gwl=j.wl, eta_ad=eta)
        dev = phbath(300.0, j.cats_l, 0.03, 200, 4.0, 256, ml=120, sig=j.SigL, gwl=j.wl, eta_ad=eta, device=cuda)
        host.gmem()
        dev.gmem(device=True)
        assert np.abs(dev.kernel - host.kernel).max() <= 1e-12 * np.abs(host.kernel).max()
        assert np.abs(np.asarray(dev.gamma) - np.asarray(host.gamma)).max() <= 1e-12 * np.abs(host.gamma).max()


def test_cosine_transform_identity(cuda):
    """A single frequency line transforms to 2 cos(w0 t) exactly (one non-zero product per row)."""
    tl = np.arange(50) * 0.7
    w = np.array([0.0, 0.3, 0.9])
    g = np.zeros((3, 2, 2)); g[1, 0, 1] = 1.5
    out = cosine_transform_device(tl, w, g, 0.0, 1.0, cuda)
    assert np.abs(out[:, 0, 1] - 3.0 * np.cos(0.3 * tl)).max() < 1e-14
    assert np.count_nonzero(out[:, 0, 0]) == 0 and np.count_nonzero(out[:, 1, :]) == 0


def junction_selfenergy(j, wl):
    """Retarded self-energy of both leads on the nd x nd system block and a Hermitian noise matrix."""
    nd = j.nd
    se = np.zeros((len(wl), nd, nd), dtype=complex)
    cl = np.array([3 * c + d for c in j.cats_l for d in range(3)])
    cr = np.array([3 * c + d for c in j.cats_r for d in range(3)])
    sl = np.array([O.flinterp(w, j.wl, j.SigL) for w in wl])
    sr = np.array([O.flinterp(w, j.wl, j.SigR) for w in wl])
    se[:, cl[:, None], cl[None, :]] += sl
    se[:, cr[:, None], cr[None, :]] += sr
    noise = -2.0 * se.imag * (1.0 + np.asarray(wl))[:, None, None] + 0j
    noise[:, cl[0], cl[1]] += 1e-6j           # make it genuinely complex Hermitian
    noise[:, cl[1], cl[0]] -= 1e-6j
    return se, noise


@pytest.mark.parametrize("nd,nc", [(12, 6), (39, 15), (60, 15), (96, 27), (117, 9)])
def test_harmonic_spectra_device_vs_oracle(cuda, nd, nc):
    """DDos and UU = Tr[D Pi D^+] of every frequency: relative error <= 1e-9 of the largest value
    (Gauss-Jordan with partial pivoting vs LAPACK getrf/getri; the lead self-energy keeps the
    matrices well conditioned)."""
    j = synth.synth(nd, nc, nc, seed=nd)
    wl = np.linspace(1e-4, 0.025, 67)
    se, noise = junction_selfenergy(j, wl)
    Ds = O.Drs(wl, j.DynMat, se)
    ddos, uu, pp = H.spectra_device(wl, j.DynMat, se, noise, cuda)
    rd, ru, rp = O.DDos(wl, Ds), O.UU(wl, Ds, noise), O.PP(wl, Ds, noise)
    assert np.abs(ddos - rd).max() <= 1e-9 * np.abs(rd).max()
    assert np.abs(uu - ru).max() <= 1e-9 * np.abs(ru).max()
    assert np.abs(pp - rp).max() <= 1e-9 * np.abs(rp).max()
    only, none_uu, none_pp = H.spectra_device(wl, j.DynMat, se, None, cuda)
    assert np.array_equal(only, ddos) and none_uu is None and none_pp is None


def test_harmonic_device_needs_pivoting(cuda):
    """A matrix with a zero leading diagonal element inverts correctly (row interchanges)."""
    rng = np.random.default_rng(0)
    n = 8
    dyn = rng.standard_normal((n, n)); dyn = dyn + dyn.T
    wl = np.array([0.5])
    dyn[0, 0] = wl[0] ** 2                     # (w^2 - dyn)[0][0] = 0 up to eta
    se = np.zeros((1, n, n), dtype=complex)
    se[0] = 1e-3j * np.eye(n); se[0, 0, 0] = 0.0
    pi = np.eye(n)[None] + 0j
    ddos, uu, _ = H.spectra_device(wl, dyn, se, pi, cuda)
    Ds = O.Drs(wl, dyn, se)
    assert np.abs(ddos - O.DDos(wl, Ds)).max() <= 1e-10 * np.abs(O.DDos(wl, Ds)).max()
    assert np.abs(uu - O.UU(wl, Ds, pi)).max() <= 1e-10 * np.abs(O.UU(wl, Ds, pi)).max()


def test_harmonic_device_size_limit(cuda):
    from pymd_b200 import _lib
    nmax = _lib.load().pymd_harmonic_max_n()
    assert nmax >= 117
    n = nmax + 1
    with pytest.raises(_lib.PymdError):
        H.spectra_device(np.array([0.1]), np.eye(n), np.zeros((1, n, n), dtype=complex), None, cuda)


# ---- batched Jacobi eigen-decomposition of the noise covariances (pymd_eigh_factors) -------------
def check_factors(fac, amat, tol=2e-13):
    """lam ascending and equal to eigvalsh; U unitary; U diag(lam) U^H = C; L L^H = U max(lam,0) U^H
    -- all relative to the Frobenius norm of each matrix."""
    lam, U, L = fac.lam, fac.U, fac.L
    ref = np.linalg.eigvalsh(amat)
    nrm = np.maximum(np.linalg.norm(amat, axis=(1, 2)), 1e-300)
    n = amat.shape[1]
    assert np.all(np.diff(lam, axis=1) >= 0)
    assert (np.abs(lam - ref).max(axis=1) / nrm).max() <= tol
    uh = np.conj(np.transpose(U, (0, 2, 1)))
    assert np.abs(uh @ U - np.eye(n)).max() <= tol
    assert (np.abs((U * lam[:, None, :]) @ uh - amat).max(axis=(1, 2)) / nrm).max() <= tol
    pos = (U * np.maximum(lam, 0)[:, None, :]) @ uh
    assert (np.abs(L @ np.conj(np.transpose(L, (0, 2, 1))) - pos).max(axis=(1, 2)) / nrm).max() <= tol
    assert fac.sweeps.max() < 40


@pytest.mark.parametrize("n,cplx", [(1, False), (2, True), (5, False), (15, False), (15, True), (39, True),
                                     (60, True), (64, False), (81, False), (83, True), (118, False),
                                     (84, True), (96, True), (118, True), (119, False), (150, False)])
def test_eigh_factors_random_matrices(cuda, n, cplx):
    """n <= 83 complex / 118 real: matrix and eigenvectors in shared memory; above (up to 118 / 167) the
    eigenvectors accumulate in a global scratch tile per CTA."""
    from pymd_b200.noise import DeviceSpectralFactors
    rng = np.random.default_rng(n)
    nf = 40
    a = rng.standard_normal((nf, n, n)) + (1j * rng.standard_normal((nf, n, n)) if cplx else 0.0)
    amat = a + np.conj(np.transpose(a, (0, 2, 1)))
    amat[0] = 0.0                                       # the covariance above the cut-off is exactly zero
    amat[1] = 1.0                                       # rank one, n-1 degenerate zeros
    amat[2] = np.diag(np.arange(n) - 2.0)               # already diagonal, negative eigenvalues
    if n > 2:
        amat[3] = amat[3] * 1e-12 + np.diag(np.repeat([1.0, 2.0], [n // 2, n - n // 2]))   # two clusters
    fac = DeviceSpectralFactors(amat, cuda, keep_vectors=True)
    assert fac.complex == cplx and fac.shape == (nf, n, n)
    check_factors(fac, amat)
    assert np.count_nonzero(fac.L[0]) == 0 and fac.sweeps[0] == 0 and fac.sweeps[2] == 0


def test_eigh_factors_of_bath_covariances(cuda):
    """The covariances the baths really decompose (config-2 shapes, a few hundred frequencies): device
    factors against LAPACK -- same eigenvalues, same L L^H."""
    from pymd_b200 import noise as PN
    j = synth.synth(60, 15, 15, seed=2)
    nmd, dt = 512, 1.0
    gam = -np.imag(j.SigL[1:]) / j.wl[1:, None, None]
    host = PN.phonon_factors(gam, j.wl[1:], 300.0, 0.03, dt, nmd)
    dev = PN.phonon_factors(gam, j.wl[1:], 300.0, 0.03, dt, nmd, device=cuda, keep_vectors=True)
    assert isinstance(dev, PN.DeviceSpectralFactors) and not dev.complex
    scale = np.abs(host.lam).max()
    assert np.abs(dev.lam - host.lam).max() <= 1e-13 * scale
    hl = host.L @ np.transpose(host.L, (0, 2, 1))
    dl = dev.L @ np.transpose(dev.L, (0, 2, 1))
    assert np.abs(dl - hl).max() <= 1e-13 * scale
    eh = PN.electron_factors(j.efric, j.exim, j.exip, 0.5, 300.0, 1.0, dt, nmd)
    ed = PN.electron_factors(j.efric, j.exim, j.exip, 0.5, 300.0, 1.0, dt, nmd, device=cuda, keep_vectors=True)
    assert ed.complex
    scale = np.abs(eh.lam).max()
    assert np.abs(ed.lam - eh.lam).max() <= 1e-13 * scale
    hl = eh.L @ np.conj(np.transpose(eh.L, (0, 2, 1)))
    dl = ed.L @ np.conj(np.transpose(ed.L, (0, 2, 1)))
    assert np.abs(dl - hl).max() <= 1e-13 * scale
    with pytest.raises(ValueError):
        PN._decompose(np.zeros((2, 130, 130), dtype=complex), cuda)
    assert isinstance(PN._decompose(np.zeros((2, 130, 130), dtype=complex), cuda, strict=False), PN.SpectralFactors)


def test_trajectory_parity_with_device_factors(cuda):
    """Whole path with set-up on the device: friction kernel from pymd_gamt, spectral factors from
    pymd_eigh_factors; the oracle gets the same (lam, U) and kernel tables, so trajectories must agree
    to the usual short-horizon tolerance."""
    import pymd_testlib as T
    c = T.CASES["eph"]
    B = 3
    m, j = T.device_md(c, B)
    for bath in m.baths:
        bath.spectral_factors()                       # host factors are cached ...
        bath.setup_device = True                      # ... and dropped when the set-up target changes
        bath.setup_keep_vectors = True
        if getattr(bath, "ml", 1) > 1:
            host_kernel = np.array(bath.kernel)
            bath.gmem()
            assert np.abs(bath.kernel - host_kernel).max() <= 1e-12 * np.abs(host_kernel).max()
    xi, r = T.draw(c, m, B, seed=11)
    m.initialise(r=r)
    m.ResetHis()
    for b, bath in enumerate(m.baths):
        bath.gnoi(xi=xi[0][b])
        assert type(bath.spectral_factors()).__name__ == "DeviceSpectralFactors"
    m.ResetSavepq()
    m.steps(c["nmd"])
    for n in (0, B - 1):
        om = T.oracle_md(c, j, share_kernels_from=m)
        om.initialise(r=r[:, n])
        om.ResetHis()
        for b, ob in enumerate(om.baths):
            ob.gnoi(xi=xi[0][b][:, :, n], factors=m.baths[b].spectral_factors().as_oracle())
        om.ResetSavepq()
        for _ in range(c["nmd"]):
            om.vv()
        assert T.relerr(m.ps[..., n], om.ps) < 1e-11 and T.relerr(m.qs[..., n], om.qs) < 1e-11
        for b, ob in enumerate(om.baths):
            assert T.relerr(m.baths[b].noise[..., n], ob.noise) < 1e-11
            assert T.relerr(m.baths[b].cur[..., n], ob.cur) < 1e-10


def test_md_power2_matches_harmonic_spectrum(cuda):
    """The always-on self-check of the reference (harmonic.py:70-84, MD power2 ~ PP): 1024 trajectories
    of the 39-dof chain between two leads, two steady-state segments of 4096 steps, against
    md.HarmonicSpectra() (self-energy rebuilt from the discrete kernel, one pymd_harmonic launch).
    Inside the phonon band the band-integrated ratio is 1 within 3 % (statistical error ~0.5 %); the
    band edge is excluded (the integrator shifts the sharp edge peak by (w dt)^2/24)."""
    import pymd_testlib as T
    c = dict(T.CASES["chain39"], nmd=4096)
    m, j = T.device_md(c, 1024, seed=3)
    m.initialise()
    m.ResetHis()
    acc = 0.0
    for rep in range(3):
        for b in m.baths:
            b.gnoi()
        m.ResetSavepq(True)
        m.steps(c["nmd"])
        if rep > 0:
            m.GetPower()
            acc = acc + m.power2[:, 1] / 2.0
    h = m.HarmonicSpectra()
    k = np.rint(h["w"] / (2 * np.pi / c["dt"] / c["nmd"])).astype(int)
    assert np.array_equal(m.power2[k, 0], h["w"])
    mdp, pp, w = acc[k], h["PP"], h["w"]
    for lo, hi in ((0.0038, 0.0075), (0.0075, 0.0112), (0.0112, 0.0148), (0.0148, 0.0184)):
        sel = (w >= lo) & (w < hi)
        assert abs(mdp[sel].sum() / pp[sel].sum() - 1.0) < 0.03, (lo, hi)
    sel = w < 0.0184
    assert abs(mdp[sel].sum() / pp[sel].sum() - 1.0) < 0.02
    assert np.all(h["DDos"][w < 0.0184] > 0)
