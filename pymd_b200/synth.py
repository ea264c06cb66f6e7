"""Synthetic junctions.

The reference's examples read SIESTA/Inelastica outputs that do not ship
(``EPH.nc``, ``Sig.nc``, ``avLambda.nc``); every config is therefore
generated: a 1-D nearest-neighbour chain (x) I_3 or a random SPD dynamical
matrix, the analytic self-energy of a semi-infinite chain for the phonon
leads, and random (anti)symmetric electron-coupling matrices with the
symmetries ``ebath.CheckEmat`` enforces (reference ebath.py:56-130).

Pure numpy, deterministic in ``seed``; used by tests, bench.py and examples.
"""
import numpy as np


class Junction:
    """Container mirroring the fields the reference reads from ``EPH.nc``
    (myio.py:146-200: Wlist, DynMat, SigL/SigR, Friction, NC, NCP, zeta1/2)."""
    pass


def chain_dynmat(na, k=1.0e-4):
    """Nearest-neighbour chain with clamped ends, K = k(2 I - S - S^T) (x) I_3."""
    a = 2.0 * k * np.eye(na)
    for i in range(na - 1):
        a[i, i + 1] = a[i + 1, i] = -k
    return np.kron(a, np.eye(3))


def random_spd_dynmat(nd, rng, lo=1.0e-5, hi=4.0e-4):
    """Q diag(lambda) Q^T with lambda ~ U(lo, hi)."""
    q, _ = np.linalg.qr(rng.standard_normal((nd, nd)))
    lam = rng.uniform(lo, hi, nd)
    return (q * lam) @ q.T


def lead_selfenergy(wl, nc, k=1.0e-4, rng=None, mix=0.3):
    """Retarded self-energy of a semi-infinite chain, Pi(w) = k^2 g(w), spread
    over ``nc`` connected dofs through a fixed SPD shape matrix so that the
    friction kernel is a genuinely matrix-valued operand.

    g(w) is the surface Green's function: inside the band (0 < w < 2 sqrt k)
    g = [(w^2-2k) - i sqrt(4k^2-(w^2-2k)^2)] / (2k^2); real above it."""
    wl = np.asarray(wl, dtype=float)
    x = wl ** 2 - 2.0 * k
    disc = 4.0 * k * k - x * x
    g = np.where(disc > 0,
                 (x - 1j * np.sqrt(np.abs(disc))) / (2.0 * k * k),
                 (x - np.sign(x) * np.sqrt(np.abs(disc))) / (2.0 * k * k) + 0j)
    if rng is None or mix == 0:
        shape = np.eye(nc)
    else:
        r = rng.standard_normal((nc, nc))
        shape = np.eye(nc) + mix * (r @ r.T) / nc
    return (k * k * g)[:, None, None] * shape[None, :, :]


def electron_matrices(nc, rng, fric=1.0e-4, nonc=2.0e-8):
    """efric SPD, exip/zeta1 symmetric, exim/zeta2 antisymmetric."""
    a = rng.standard_normal((nc, nc))
    b = rng.standard_normal((nc, nc))
    c = rng.standard_normal((nc, nc))
    d = rng.standard_normal((nc, nc))
    efric = fric * (a @ a.T / nc + 0.5 * np.eye(nc))
    exip = nonc * (b + b.T)
    exim = nonc * (b - b.T)
    zeta1 = nonc * (c + c.T)
    zeta2 = nonc * (d - d.T)
    return efric, exim, exip, zeta1, zeta2


def synth(nd, nc_l, nc_r, seed=0, chain=True, k=1.0e-4, ngw=601, wtop=0.06,
          nc_e=None, constraint=False):
    """Synthetic junction: ``nd`` system dofs (multiple of 3), left/right phonon
    leads on the first ``nc_l`` / last ``nc_r`` dofs, optional electron bath on
    the first ``nc_e`` dofs, optional single constraint vector (centre-of-mass
    direction, as examples/AuBZ.py:154-169 builds one)."""
    assert nd % 3 == 0 and nc_l % 3 == 0 and nc_r % 3 == 0
    rng = np.random.default_rng(seed)
    j = Junction()
    j.nd, j.na = nd, nd // 3
    j.DynMat = chain_dynmat(j.na, k) if chain else random_spd_dynmat(nd, rng)
    j.wl = np.linspace(0.0, wtop, ngw)
    j.SigL = lead_selfenergy(j.wl, nc_l, k, rng)
    j.SigR = lead_selfenergy(j.wl, nc_r, k, rng)
    j.cats_l = list(range(nc_l // 3))
    j.cats_r = list(range(j.na - nc_r // 3, j.na))
    j.axyz = [["Au", 0.0, 0.0, 2.5 * i] for i in range(j.na)]
    if nc_e is None:
        nc_e = nd
    j.cats_e = list(range(nc_e // 3))
    erng = np.random.default_rng(seed + 1)
    j.efric, j.exim, j.exip, j.zeta1, j.zeta2 = electron_matrices(nc_e, erng)
    if constraint:
        c = np.zeros(nd)
        c[2::3] = 1.0
        j.constr = np.array([c])
    else:
        j.constr = None
    return j
