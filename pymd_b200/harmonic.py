"""Analytic harmonic spectra from retarded Green's functions (reference harmonic.py:21-96):
the statistical cross-check for MD spectra (MD power2 ~ PP).  Host numpy, batched over
frequency, plus ``spectra_device`` which inverts every frequency's matrix in one GPU launch
(csrc/setup.cu).  Not on the stepping path."""
import numpy as N

from . import _lib
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


def spectra_device(wl, dyn, se, noise=None, device=None, eta=1e-10):
    """(DDos, UU, PP) of Drs/DDos/UU/PP above computed by libpymd_b200 (pymd_harmonic): one CTA
    inverts (w + i eta)^2 - dyn - se(w) per frequency in shared memory.  ``se`` and ``noise`` are
    (nw, n, n) complex; without ``noise`` UU and PP are None.  n <= pymd_harmonic_max_n()."""
    import torch
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    wl = N.ascontiguousarray(wl, dtype=float)
    dyn = N.ascontiguousarray(dyn, dtype=float)
    n, nw = len(dyn), len(wl)
    chkShape(dyn)
    se = N.array(N.broadcast_to(N.asarray(se, dtype=complex), (nw, n, n)))
    wl_d = torch.as_tensor(wl, device=device)
    dyn_d = torch.as_tensor(dyn, device=device)
    se_d = torch.as_tensor(se, device=device)
    pi_d = uu_d = None
    if noise is not None:
        pi_d = torch.as_tensor(N.array(N.broadcast_to(N.asarray(noise, dtype=complex), (nw, n, n))), device=device)
        uu_d = torch.empty(nw, dtype=torch.float64, device=device)
    ddos_d = torch.empty(nw, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        _lib.call("pymd_harmonic", _lib.ptr(wl_d), nw, _lib.ptr(dyn_d), _lib.ptr(se_d), _lib.ptr(pi_d), n,
                  float(eta), _lib.ptr(ddos_d), _lib.ptr(uu_d), _lib.stream_ptr())
    ddos = ddos_d.cpu().numpy()
    if noise is None:
        return ddos, None, None
    uu = uu_d.cpu().numpy()
    return ddos, uu, wl ** 2 * uu
