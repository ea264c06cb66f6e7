"""Analytic harmonic spectra from retarded Green's functions (reference harmonic.py:21-96):
the statistical cross-check for MD spectra (MD power2 ~ PP).  Host numpy; batched over
frequency.  Not on the stepping path."""
import numpy as N

from .matrix import chkShape


def Dr(w, dyn, se):
    """D^r(w) = [(w + i eta)^2 - K - Sigma(w)]^-1, eta = 1e-10 (harmonic.py:37-40)."""
    n = len(dyn)
    return N.linalg.inv((w + 1.j * 10 ** -10) ** 2 * N.identity(n) - dyn - se)


def Drs(wl, dyn, se):
    wl = N.asarray(wl)
    n = len(dyn)
    a = ((wl + 1.j * 10 ** -10) ** 2)[:, None, None] * N.identity(n)[None] - N.asarray(dyn)[None] - N.asarray(se)
    return N.linalg.inv(a)


def DDos(wl, Ds):
    """-2 w Tr Im D (harmonic.py:42-52)."""
    chkShape(Ds[0])
    return -2. * N.asarray(wl) * N.trace(N.asarray(Ds).imag, axis1=1, axis2=2)


def UU(wl, Ds, noise):
    """Tr[D Pi D^dagger] (harmonic.py:54-68)."""
    Ds, noise = N.asarray(Ds), N.asarray(noise)
    return N.einsum("wij,wjk,wik->w", Ds, noise, Ds.conj()).real


def PP(wl, Ds, noise):
    """w^2 Tr[D Pi D^dagger] (harmonic.py:70-84)."""
    return N.asarray(wl) ** 2 * UU(wl, Ds, noise)


def SigE(wl, efric, exim, zeta1, zeta2, bias):
    """Effective electronic retarded self-energy (harmonic.py:86-96)."""
    chkShape(efric)
    w = N.asarray(wl)[:, None, None]
    return -1.j * w * (efric + bias * zeta2) + bias * zeta1 - bias * exim
