"""CPU oracle for the generalized-Langevin hot path of phyjtlu/pymd.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import this module; nothing under
``pymd_b200/`` does (the product path has no CPU fallback).

This is a Python-3 / numpy restatement of the reference's algorithm for the
path named in BASELINE.json (velocity-Verlet loop, dynamical-matrix force,
bath friction with memory, coloured quantum noise, power spectra).  Every
function cites the reference file:line it follows.  It keeps the reference's
*sequence of floating-point operations* (including the three full memory
convolutions per step and the Python loop over the memory length), so that

  * with the legacy numpy global RNG seeded identically it reproduces the
    mechanically converted reference (``oracle/_ref``, see build_ref.py)
    BIT-FOR-BIT -- this is how the oracle is pinned
    (tests/test_oracle_pin.py, tests/golden/*.npz), and
  * timed on host cores it is a fair stand-in for the reference CPU path.

Parity status: the reference ships no golden vectors, KATs or fixtures for this
path.  The oracle is therefore pinned against outputs of
the reference itself, run in the build container through ``oracle/_ref`` and
committed as ``tests/golden/*.npz`` by ``tests/golden/make_golden.py``.

Two surgical additions (both off by default):
  (a) RNG injection: ``xi`` (standard normals, indexed [freq, eigen-index]) for
      the noise, ``r`` (uniforms, one per mode) for the initial condition.
      ``xi=None`` / ``r=None`` draws from ``numpy.random`` in exactly the
      reference's consumption order (no draw where an eigenvalue is <= 0).
  (b) pre-computed spectral factors may be handed in, so that the device path
      and the oracle share the *same* eigenvector gauge.
"""
import sys

import numpy as N
from numpy import linalg as LA

# ---------------------------------------------------------------- units.py:7-12
hbar = 1.0
kb = 0.000086173423
curcof = 243414.


# ------------------------------------------------------------------ matrix.py
def chkShape(a):
    """matrix.py:7-17 -- n for an n x n matrix."""
    sh = N.shape(N.array(a))
    if sh[0] != sh[1]:
        raise ValueError("The matrix should be a n by n matrix")
    return sh[0]


def symmetrize(a):
    """matrix.py:19-21."""
    aa = N.array(a)
    return 0.5 * (aa + N.transpose(aa))


def antisymmetrize(a):
    """matrix.py:23-25."""
    aa = N.array(a)
    return 0.5 * (aa - N.transpose(aa))


def dagger(a):
    """matrix.py:27-33."""
    return N.transpose(N.conjugate(N.array(a)))


def hermitianize(a):
    """matrix.py:35-37."""
    aa = N.array(a)
    return 0.5 * (aa + dagger(aa))


def mdot(*args):
    """matrix.py:39-44 -- chained dot that starts from an identity."""
    tmp = N.identity(len(args[0]))
    for a in args:
        tmp = N.dot(tmp, a)
    return tmp


def mm(*args):
    """functions.py:97-101 / matrix.py:46-51 -- chained dot."""
    tmp = args[0].copy()
    for a in args[1:]:
        tmp = N.dot(tmp, a)
    return tmp


# --------------------------------------------------------------- functions.py
def bose(w, T):
    """functions.py:26-45 -- Bose-Einstein with the reference's special cases."""
    if T == 0.0:
        if w == 0.0:
            return 1 / (N.exp(1.0 / kb) - 1)
        elif w < 0.0:
            return -1.0
        return 0.0
    if w == 0.0:
        return 0.0
    return 1.0 / (N.exp(w / kb / T) - 1.0)


def nearest(b, bs):
    """functions.py:80-86 -- index of the first element of bs closest to b."""
    d = abs(N.array(bs) - b)
    return list(d).index(min(d))


def flinterp(x, xs, ys):
    """functions.py:61-78 -- 'linear interpolation' including the reference's
    sign convention in the slope term (value mirrored about the nearest grid
    point; exact on grid points)."""
    i = nearest(x, xs)
    if i == len(xs) - 1:
        return ys[-1]
    if i == 0:
        return ys[0]
    dd = xs[i] - x
    if dd > 0:
        return ys[i] + dd / (xs[i] - xs[i - 1]) * (ys[i] - ys[i - 1])
    return ys[i] + dd / (xs[i] - xs[i + 1]) * (ys[i] - ys[i + 1])


def rpadleft(bs, b):
    """functions.py:88-95 -- shift a new row in at the front, drop the last."""
    if len(bs) > 1:
        return N.concatenate((N.array([b]), N.array(bs)[:-1]), axis=0)
    if len(bs) == 1:
        return N.array([b])
    raise ValueError("len(bs) is less than 1")


# ------------------------------------------------------------------- myfft.py
def Fourier1D(a, dt):
    """myfft.py:15-32 -- t->w: ifft(a) * (2 pi / dw),  dw = 2 pi/dt/n."""
    n = len(a)
    dw = 2 * N.pi / dt / n
    return (2. * N.pi / dw) * N.fft.ifft(a)


def iFourier1D(a, dt):
    """myfft.py:35-52 -- w->t: fft(a) * (dw/2/pi)."""
    n = len(a)
    dw = 2 * N.pi / dt / n
    return (dw / 2 / N.pi) * N.fft.fft(a)


# ------------------------------------------------------------------- noise.py
def equ(w, cut, T, classical=False, zpmotion=True):
    """noise.py:229-250 -- 2 hw (n_B + zp) below the cutoff, with the
    classical / w==0 limits 2 kb T."""
    hw = hbar * w
    zp = 0.5 if zpmotion is True else 0.0
    if hw < cut:
        if classical:
            return 2.0 * kb * T
        if hw == 0:
            return 2.0 * kb * T
        return 2.0 * hw * (zp + bose(hw, T))
    return 0.0


class _Normals:
    """Source of standard normals in the reference's consumption order."""

    def __init__(self, xi=None):
        self.xi = xi

    def get(self, f, i):
        if self.xi is None:
            return N.random.standard_normal()
        return self.xi[f, i]


def vargau(av, au, f, src):
    """noise.py:252-284 -- r_i ~ N(0, sqrt(av_i)) for av_i > 0 (no draw
    otherwise), returned as au . r.  ``normal(0, s)`` is ``s * standard_normal()``
    for the same generator state, which is what makes injection exact."""
    rval = []
    for i in range(len(av)):
        if av[i] <= 0:
            rval.append(0.0)
        else:
            rval.append(N.sqrt(av[i]) * src.get(f, i))
    return mdot(N.array(au), rval)


def phonon_cov(i, gamma, wl, T, phcut, dt, nmd, classical, zpmotion):
    """noise.py:61-66 -- covariance of the phonon noise at frequency index i."""
    dw = 2.0 * N.pi / dt / nmd
    delta = dt * nmd
    w = dw * i
    amat = delta * equ(w, phcut, T, classical, zpmotion) * flinterp(w, wl, gamma)
    return hermitianize(amat)


def electron_cov(i, efric, exim, exip, bias, T, ecut, dt, nmd, classical, zpmotion):
    """noise.py:156-170 -- covariance of the electron noise at frequency index i
    (equilibrium part + the two bias-shifted non-equilibrium parts)."""
    dw = 2.0 * N.pi / dt / nmd
    delta = dt * nmd
    w = dw * i
    aw = delta * equ(w, ecut, T, classical, zpmotion)
    amate = aw * N.array(efric)
    awm = delta * equ(hbar * w - bias, ecut, T, classical, zpmotion)
    amatm = -0.5 * aw * N.array(exip) + 0.5 * awm * (N.array(exip) + 1j * N.array(exim))
    awp = delta * equ(hbar * w + bias, ecut, T, classical, zpmotion)
    amatp = -0.5 * aw * N.array(exip) + 0.5 * awp * (N.array(exip) - 1j * N.array(exim))
    return hermitianize(amate + amatm + amatp)


def _mirror_and_transform(half, dt, nmd):
    """noise.py:74-88 / 176-190 -- Hermitian mirror of the half spectrum (the
    Nyquist row is replaced by its conjugate), then w->t column by column."""
    hlen = int(nmd / 2)
    half = N.array(half)
    neg = N.array([N.conjugate(half[hlen - i, :]) for i in range(hlen)])
    full = N.concatenate((N.delete(half, -1, 0), neg), axis=0)
    cols = [iFourier1D(full[:, i], dt) for i in range(full.shape[1])]
    return N.transpose(N.array(cols))


def phnoise(gamma, wl, T, phcut, dt, nmd, classical=False, zpmotion=True,
            xi=None, factors=None):
    """noise.py:38-88 -- coloured phonon noise, shape (nmd, nc), complex.
    ``factors=(lam, U)`` bypasses eigh so both sides share one gauge."""
    hlen = int(nmd / 2)
    src = _Normals(xi)
    half = []
    for i in range(hlen + 1):
        if factors is None:
            av, au = LA.eigh(phonon_cov(i, gamma, wl, T, phcut, dt, nmd, classical, zpmotion))
        else:
            av, au = factors[0][i], factors[1][i]
        half.append(vargau(av, au, i, src))
    return _mirror_and_transform(half, dt, nmd)


def enoise(efric, exim, exip, bias, T, ecut, dt, nmd, classical=False, zpmotion=True,
           xi=None, factors=None):
    """noise.py:135-190 -- coloured electron noise, shape (nmd, nc), complex."""
    hlen = int(nmd / 2)
    src = _Normals(xi)
    half = []
    for i in range(hlen + 1):
        if factors is None:
            av, au = LA.eigh(electron_cov(i, efric, exim, exip, bias, T, ecut, dt, nmd,
                                          classical, zpmotion))
        else:
            av, au = factors[0][i], factors[1][i]
        half.append(vargau(av, au, i, src))
    return _mirror_and_transform(half, dt, nmd)


def phnoisew(gamma, wl, T, phcut, classical=False, zpmotion=True):
    """noise.py:16-34 -- phonon noise spectrum on the wl grid, no sampling."""
    gamma = N.array(gamma)
    out = 0.0 * gamma
    for i in range(len(wl)):
        out[i] = equ(wl[i], phcut, T, classical, zpmotion) * gamma[i]
    return out


def enoisew(wl, efric, exim, exip, bias, T, ecut, classical=False, zpmotion=True):
    """noise.py:91-129 -- electron noise spectrum on the wl grid, no sampling."""
    nc = chkShape(efric)
    out = N.zeros((len(wl), nc, nc), complex)
    for i in range(len(wl)):
        w = wl[i]
        aw = equ(w, ecut, T, classical, zpmotion)
        awm = equ(hbar * w - bias, ecut, T, classical, zpmotion)
        awp = equ(hbar * w + bias, ecut, T, classical, zpmotion)
        amate = aw * N.array(efric)
        amatm = -0.5 * aw * N.array(exip) + 0.5 * awm * (N.array(exip) + 1j * N.array(exim))
        amatp = -0.5 * aw * N.array(exip) + 0.5 * awp * (N.array(exip) - 1j * N.array(exim))
        out[i] = hermitianize(amate + amatm + amatp)
    return out


# ---------------------------------------------------------------- spectrum.py
def powerspec(qs, dt, nmd):
    """spectrum.py:8-25 -- w^2 sum_dof |q(w)|^2 /(dt nmd), rows (w_i, value)."""
    qst = N.transpose(N.array(qs))
    if qst.shape[1] != nmd:
        raise ValueError("power: qs shape error!")
    dw = 2. * N.pi / dt / nmd
    qsw = N.array([Fourier1D(a, dt) for a in qst])
    qsw = N.real(N.transpose(qsw * N.conjugate(qsw)))
    return N.array([[i * dw, (dw * i) ** 2 * N.sum(qsw[i]) / dt / nmd] for i in range(nmd)])


def powerspec2(ps, dt, nmd):
    """spectrum.py:28-45 -- sum_dof |p(w)|^2 /(dt nmd), rows (w_i, value)."""
    pst = N.transpose(N.array(ps))
    if pst.shape[1] != nmd:
        raise ValueError("power: ps shape error!")
    dw = 2. * N.pi / dt / nmd
    psw = N.array([Fourier1D(a, dt) for a in pst])
    psw = N.real(N.transpose(psw * N.conjugate(psw)))
    return N.array([[i * dw, N.sum(psw[i]) / dt / nmd] for i in range(nmd)])


# ------------------------------------------------------------------ phbath.py
def mf(f, cids, lens):
    """phbath.py:207-214 -- scatter the nc-vector f into an nd-vector of zeros."""
    t = N.zeros(lens)
    for i in range(len(cids)):
        t[cids[i]] = f[i]
    return t


def exlist(a, indices):
    """phbath.py:217-218 -- gather."""
    return N.array([a[i] for i in indices])


def gamt(tl, wl, gwl, gam, eta_ad=0):
    """phbath.py:221-254 -- friction kernel in time: cosine transform of the
    (mirror-interpolated) gamma(w) on the wl grid; eta_ad != 0 adds the
    artificial-damping factors."""
    gt = []
    if eta_ad == 0:
        for t in tl:
            tm = [N.array(flinterp(wl[i], gwl, gam)) * N.cos(wl[i] * t) for i in range(len(wl))]
            gt.append(2.0 * N.mean(N.array(tm), axis=0) * wl[-1] / N.pi)
    else:
        for t in tl:
            tm = []
            for i in range(len(wl)):
                g = N.array(flinterp(wl[i], gwl, gam))
                tm.append(g * wl[i] / (wl[i] - 1j * eta_ad) * N.exp(-1j * wl[i] * t - eta_ad * t)
                          + g * wl[i] / (wl[i] + 1j * eta_ad) * N.exp(+1j * wl[i] * t - eta_ad * t))
            gt.append(N.mean(N.array(tm), axis=0) * wl[-1] / N.pi)
    return N.array(N.real(gt))


class phbath:
    """phbath.py:9-202 -- phonon bath (same constructor as the reference)."""

    def __init__(self, T, cats, debye, nw, dt, nmd, ml=None, mcof=2.0, sig=None, gamma=None,
                 gwl=None, K00=None, K01=None, V01=None, eta_ad=0, classical=False,
                 zpmotion=True):
        self.classical, self.zpmotion = classical, zpmotion
        self.T, self.debye, self.cats = T, debye, N.array(cats, dtype='int')
        self.dt, self.nmd, self.ml = dt, nmd, ml
        self.kernel = None
        self.cids = N.array([[3 * c + 0, 3 * c + 1, 3 * c + 2] for c in cats]).flatten()
        self.nc = len(self.cids)
        self.wmax = mcof * debye
        self.local = False
        self.nw = nw
        self.wl = [self.wmax * i / nw for i in range(nw)]
        self.gamma, self.sig, self.gwl = gamma, sig, gwl
        self.cur = N.zeros(nmd)
        self.eta_ad = eta_ad
        self.noise = None
        if K00 is not None and K01 is not None and V01 is not None:      # phbath.py:66-70
            raise NotImplementedError("phbath: Calculating self-energy is not implemented yet.")
        elif (sig is not None or gamma is not None) and gwl is not None:  # phbath.py:72-81
            if sig is not None:
                if len(self.sig[0]) != self.nc:
                    raise ValueError("phbath: inconsist between cids and sig!")
                self.ggamma()
            if self.gamma is not None and len(self.gamma[0]) != self.nc:
                raise ValueError("phbath: inconsist between cids and gamma!")
        else:                                                             # phbath.py:82-89
            phfric = debye * N.pi / 6.0
            self.gamma = N.array([N.diag(phfric + N.zeros(int(self.nc)))])
            self.gwl = N.array([0])
            self.local = True
            self.ml = 1

    def SetT(self, T):
        self.T = T

    def ggamma(self):
        """phbath.py:126-144 -- gamma = -Im Sig / w, with the w==0 entry taken
        from the next grid point."""
        a = []
        for i in range(len(self.gwl)):
            if self.gwl[i] == 0:
                a.append(-N.imag(self.sig[i + 1]) / self.gwl[i + 1])
            else:
                a.append(-N.imag(self.sig[i]) / self.gwl[i])
        self.gamma = N.array(a)

    def gnoi(self, xi=None, factors=None):
        """phbath.py:146-158."""
        self.noise = N.real(phnoise(self.gamma, self.gwl, self.T, self.wmax, self.dt, self.nmd,
                                    self.classical, self.zpmotion, xi=xi, factors=factors))

    def gmem(self):
        """phbath.py:161-190."""
        if self.local:
            self.ml = 1
            self.kernel = self.gamma
            return
        tl = [self.dt * i for i in range(self.ml)]
        self.kernel = N.real(gamt(tl, self.wl, self.gwl, self.gamma, self.eta_ad))
        if self.eta_ad != 0:                                              # phbath.py:175-189
            gamnew = N.zeros(N.shape(self.gamma))
            for i in range(len(self.gwl)):
                for it in range(self.kernel.shape[0]):
                    gamnew[i, :, :] = gamnew[i, :, :] + self.dt * self.kernel[it, :, :] * N.cos(self.gwl[i] * tl[it])
            self.gammaOld = self.gamma
            self.gamma = N.array(N.real(gamnew))

    def bforce(self, t, phis, qhis):
        """phbath.py:192-202 -- noise minus the memory-kernel friction."""
        f = self.noise[t % self.nmd]
        for i in range(self.ml):
            if self.ml == 1:
                f = f - mm(self.kernel[i], exlist(phis[i], self.cids))
            else:
                f = f - mm(self.kernel[i], exlist(phis[i], self.cids)) * self.dt
        return mf(f, self.cids, len(phis[0]))


# ------------------------------------------------------------------- ebath.py
class ebath:
    """ebath.py:10-199 -- electron bath (time-local friction, ml = 1)."""

    def __init__(self, cats, T, dt, nmd, wmax=None, nw=None, bias=0., efric=None, exim=None,
                 exip=None, zeta1=None, zeta2=None, classical=False, zpmotion=True):
        self.cats = N.array(cats, dtype='int')
        self.cids = N.array([[3 * c + 0, 3 * c + 1, 3 * c + 2] for c in cats]).flatten()
        self.nc = len(self.cids)
        self.T, self.wmax, self.nw, self.bias = T, wmax, nw, bias
        self.dt, self.nmd = dt, nmd
        self.cur = N.zeros(nmd)
        self.classical, self.zpmotion = classical, zpmotion
        self.wl = None if (nw is None or wmax is None) else [wmax * i / nw for i in range(nw)]
        self.CheckEmat(efric, exim, exip, zeta1, zeta2)
        self.ml = 1
        self.noise = None

    def CheckEmat(self, efric=None, exim=None, exip=None, zeta1=None, zeta2=None):
        """ebath.py:56-130 -- (anti)symmetrise the five coupling matrices."""
        if efric is None:
            self.efric = self.kernel = self.exim = self.exip = self.zeta1 = self.zeta2 = None
            self.ebath = False
            return
        n = chkShape(efric)
        if n != self.nc:
            raise ValueError("ebath.CheckEmat: efric shape error!")
        self.efric = symmetrize(efric)
        self.kernel = N.array([self.efric])
        self.exip = N.zeros(shape=(n, n))
        self.exim = N.zeros(shape=(n, n))
        self.zeta1 = N.zeros(shape=(n, n))
        self.zeta2 = N.zeros(shape=(n, n))
        self.ebath = True
        for name, m, op in (("exim", exim, antisymmetrize), ("exip", exip, symmetrize),
                            ("zeta1", zeta1, symmetrize), ("zeta2", zeta2, antisymmetrize)):
            if m is not None:
                if chkShape(m) != self.nc:
                    raise ValueError("ebath.CheckEmat: the dimension of %s is wrong!" % name)
                setattr(self, name, op(m))

    def gnoi(self, xi=None, factors=None):
        """ebath.py:132-148."""
        if self.ebath is False:
            raise ValueError("ebath.gnoi: ebath is False!")
        self.noise = N.real(enoise(self.efric, self.exim, self.exip, self.bias, self.T, self.wmax,
                                   self.dt, self.nmd, self.classical, self.zpmotion,
                                   xi=xi, factors=factors))

    def GetSig(self):
        """ebath.py:150-165 -- wide-band effective retarded self-energy."""
        nc = chkShape(self.efric)
        self.sig = N.zeros((len(self.wl), nc, nc), complex)
        for i in range(len(self.wl)):
            self.sig[i] = -1.j * self.wl[i] * (self.efric + self.bias * self.zeta2) \
                + self.bias * self.zeta1 - self.bias * self.exim

    def setbias(self, bias=0.0):
        self.bias = bias

    def bforce(self, t, phis, qhis):
        """ebath.py:180-199 -- noise, friction, non-conservative, renormalisation
        and Berry terms; only the newest history rows are used."""
        f = self.noise[t % self.nmd]
        for i in range(self.ml):
            f = f - mm(self.kernel[i], exlist(phis[i], self.cids)) \
                + mm(self.bias * self.exim, exlist(qhis[0], self.cids)) \
                - mm(self.bias * self.zeta1, exlist(qhis[0], self.cids)) \
                - mm(self.bias * self.zeta2, exlist(phis[0], self.cids))
        return mf(f, self.cids, len(phis[0]))


# ---------------------------------------------------------------------- md.py
def sameq(q1, q2):
    """md.py:724-736."""
    if len(q1) != len(q2):
        return False
    return bool(max(abs(N.array(q1) - N.array(q2))) < 10e-10)


def ApplyConstraint(f, constr=None):
    """md.py:739-762 -- remove the components of f along each (normalised)
    constraint vector; all dots use the un-projected f."""
    if constr is None:
        return f
    f = N.array(f)
    constr = N.array(constr)
    nf = 1.0 * f
    for i in range(N.shape(constr)[0]):
        nn = LA.norm(constr[i])
        if nn != 0:
            constr[i] = constr[i] / nn
        nf = nf - N.dot(f, constr[i]) * constr[i]
    return nf


class md:
    """md.py:23-705 -- the Langevin integrator (harmonic driver only; the
    SIESTA / Brenner / LAMMPS slots are out of scope)."""

    def __init__(self, dt, nmd, T, syslist=None, axyz=None, harmonic=False, dyn=None,
                 savepq=True, nrep=1, npie=8, constr=None):
        self.constraint = constr
        self.nrep, self.dt, self.nmd, self.harmonic, self.T, self.npie = nrep, dt, nmd, harmonic, T, npie
        self.power = N.zeros((nmd, 2))
        self.power2 = N.zeros((nmd, 2))
        self.nph = None
        if syslist is not None:
            self.syslist = N.array(syslist, dtype='int')
            self.na, self.nph = len(syslist), 3 * len(syslist)
        elif axyz is not None:
            self.syslist = N.array(range(len(axyz)), dtype='int')
            self.na, self.nph = len(axyz), 3 * len(axyz)
        self.ml, self.t = 1, 0
        self.p, self.q, self.q0, self.f0 = [], [], [], []
        self.baths, self.fhis, self.fbaths = [], [], []
        self.etot = N.zeros(nmd)
        self.setDyn(dyn)
        self.ResetSavepq(savepq)

    def setDyn(self, dyn=None):
        """md.py:211-250 -- symmetrise, clamp negative eigenvalues, rebuild."""
        if dyn is None:
            self.dyn = self.hw = self.U = None
            return
        ndyn = N.array(dyn)
        n = chkShape(ndyn)
        if self.nph is not None and self.nph != n:
            raise ValueError("md.setDyn: the dimension of dynamical matrix is wrong!")
        self.nph = n
        self.dyn = symmetrize(ndyn)
        av, au = LA.eigh(self.dyn)
        if min(av) < 0:
            avn = 0. * av
            for i in range(len(av)):
                avn[i] = 0. if av[i] < 0 else av[i]
            av = avn
        self.hw = N.array(list(map(N.real, map(N.sqrt, av))))
        self.U = N.array(au)
        self.dyn = mm(self.U, N.diag(N.array(av)), N.transpose(self.U))

    def ResetSavepq(self, savepq=True):
        """md.py:143-150."""
        self.savepq = savepq
        self.ps = N.zeros((self.nmd, self.nph))
        self.qs = N.zeros((self.nmd, self.nph))

    def energy(self):
        """md.py:154-158 (second definition wins): kinetic energy only."""
        return 0.5 * mm(self.p, self.p)

    def AddBath(self, bath):
        """md.py:160-176."""
        if self.dt != bath.dt:
            raise ValueError("md.AddBath: md time step dt not consistent!")
        if self.nmd != bath.nmd:
            raise ValueError("md.AddBath: number of md steps nmd not consistent!")
        self.baths.append(bath)
        if bath.ml > self.ml:
            self.ml = bath.ml
        self.fbaths.append(N.zeros(self.nph))
        self.fhis.append(N.zeros((self.nmd, self.nph)))

    def initialise(self, r=None):
        """md.py:253-290 -- quantum initial condition mode by mode; one uniform
        per mode is consumed even for the frozen slow modes."""
        self.t = 0
        if self.dyn is None:
            self.p = N.zeros(self.nph)
            self.q = N.zeros(self.nph)
            return
        av, au = self.hw, self.U
        dis = N.zeros(len(av))
        vel = N.zeros(len(av))
        for i in range(len(av)):
            if av[i] < 0.01:
                am = 0.0
            else:
                am = ((bose(av[i], self.T) + 0.5) * 2.0 / av[i]) ** 0.5
            ri = N.random.rand() if r is None else r[i]
            dis = dis + au[:, i] * am * N.cos(2. * N.pi * ri)
            vel = vel - av[i] * au[:, i] * am * N.sin(2. * N.pi * ri)
            dis = ApplyConstraint(dis, self.constraint)
            vel = ApplyConstraint(vel, self.constraint)
        self.p, self.q = vel, dis
        self.pinit, self.qinit = vel, dis

    def ResetHis(self):
        """md.py:292-301."""
        self.qhis = N.zeros((self.ml, self.nph))
        self.phis = N.zeros((self.ml, self.nph))

    def GetPower(self):
        """md.py:303-313."""
        self.power = powerspec(self.qs, self.dt, self.nmd)
        self.power2 = powerspec2(self.ps, self.dt, self.nmd)

    def potforce(self, q):
        """md.py:444-492, dynamical-matrix branch, with the sameq cache."""
        if sameq(q, self.q0):
            return self.f0
        f = -mdot(self.dyn, q)
        self.q0, self.f0 = q, f
        return f

    def force(self, t, p, q, id=0):
        """md.py:360-383 -- potential force plus every bath's force; for id=1
        the trial (p, q) is pushed onto a copy of the history."""
        it = t + id
        pf = self.potforce(q)
        if id == 0:
            tphis, tqhis = self.phis, self.qhis
        else:
            tphis, tqhis = rpadleft(self.phis, p), rpadleft(self.qhis, q)
        for i in range(len(self.baths)):
            self.fbaths[i] = self.baths[i].bforce(it, tphis, tqhis)
            pf = pf + self.fbaths[i]
        return pf

    def vv(self):
        """md.py:315-358 -- modified velocity-Verlet with one predictor and one
        corrector for the velocity-dependent bath force."""
        t, p, q = int(self.t), self.p, self.q
        if self.savepq:
            self.ps[t % self.nmd] = p
            self.qs[t % self.nmd] = q
        self.etot[t % self.nmd] = self.energy()
        self.qhis = rpadleft(self.qhis, q)
        self.phis = rpadleft(self.phis, p)
        f = self.force(t, p, q, 0)
        pthalf = p + f * self.dt / 2.0
        qtt = q + p * self.dt + f * self.dt ** 2 / 2.0
        for i in range(len(self.baths)):
            self.baths[i].cur[t % self.nmd] = mm(self.fbaths[i], p)
            self.fhis[i][t % self.nmd] = self.fbaths[i]
        f = self.force(t, pthalf, qtt, 1)
        ptt1 = pthalf + self.dt * f / 2.0
        f = self.force(t, ptt1, qtt, 1)
        ptt2 = pthalf + self.dt * f / 2.0
        ptt2 = ApplyConstraint(ptt2, self.constraint)
        qtt = ApplyConstraint(qtt, self.constraint)
        self.t, self.p, self.q = t + 1, ptt2, qtt

    def Run(self, r=None, xi=None, factors=None):
        """md.py:523-652 without the NetCDF checkpoints and text files: nrep
        consecutive segments, fresh noise per segment, running mean of spectra.
        ``xi[j][b]`` / ``factors[b]`` inject the white noise per segment/bath."""
        self.initialise(r)
        self.ResetHis()
        self.curmean = N.zeros((self.nrep, len(self.baths)))
        for j in range(self.nrep):
            for b, bath in enumerate(self.baths):
                bath.gnoi(xi=None if xi is None else xi[j][b],
                          factors=None if factors is None else factors[b])
            self.ResetSavepq()
            for _ in range(self.nmd):
                self.vv()
            power, power2 = N.copy(self.power), N.copy(self.power2)
            self.GetPower()
            self.power = (power * j + self.power) / float(j + 1)
            self.power2 = (power2 * j + self.power2) / float(j + 1)
            for b, bath in enumerate(self.baths):
                self.curmean[j, b] = N.mean(bath.cur) * curcof          # md.py:619


# ---------------------------------------------------------------- harmonic.py
def Dr(w, dyn, se):
    """harmonic.py:37-40 -- retarded Green's function."""
    n = len(dyn)
    return LA.inv((w + 1.j * 10 ** -10) ** 2 * N.identity(n) - dyn - se)


def Drs(wl, dyn, se):
    """harmonic.py:21-36."""
    return N.array([Dr(wl[i], dyn, se[i]) for i in range(len(wl))])


def DDos(wl, Ds):
    """harmonic.py:42-52."""
    return N.array([-2. * wl[i] * N.trace(Ds[i].imag) for i in range(len(wl))])


def UU(wl, Ds, noise):
    """harmonic.py:54-68 -- Tr[D Pi D^dagger]."""
    return N.array([N.trace(mm(Ds[i], noise[i], dagger(Ds[i]))).real for i in range(len(Ds))])


def PP(wl, Ds, noise):
    """harmonic.py:70-84 -- w^2 Tr[D Pi D^dagger]."""
    return N.array([wl[i] ** 2 * N.trace(mm(Ds[i], noise[i], dagger(Ds[i]))).real
                    for i in range(len(Ds))])


def SigE(wl, efric, exim, zeta1, zeta2, bias):
    """harmonic.py:86-96 -- effective electronic retarded self-energy."""
    nc = chkShape(efric)
    Sig = N.zeros((len(wl), nc, nc), complex)
    for i in range(len(wl)):
        Sig[i] = -1.j * wl[i] * (efric + bias * zeta2) + bias * zeta1 - bias * exim
    return Sig
