"""Phonon bath with the reference's class interface (phbath.py:9-202).

Setup tables (gamma(w), the friction kernel K(t_k)) are computed on the host with the
reference's formulas by default (bit-identical tables); ``gmem(device=True)`` or the
environment variable PYMD_SETUP_DEVICE=1 runs the cosine transform on the GPU
(csrc/setup.cu) for long kernels.  Noise generation and the memory-kernel friction inside
the time loop run on the GPU for the whole ensemble.
"""
import numpy as N

from . import _lib
from . import noise as _noise
from ._bathbase import BathBase
from .functions import flinterp


class phbath(BathBase):
    """phbath(T, cats, debye, nw, dt, nmd, ml=None, mcof=2.0, sig=None, gamma=None, gwl=None,
    K00=None, K01=None, V01=None, eta_ad=0, classical=False, zpmotion=True)  -- phbath.py:44.
    Optional ensemble knobs (not in the reference): batch, seed, device."""

    def __init__(self, T, cats, debye, nw, dt, nmd, ml=None, mcof=2.0, sig=None, gamma=None, gwl=None,
                 K00=None, K01=None, V01=None, eta_ad=0, classical=False, zpmotion=True,
                 batch=None, seed=None, device=None):
        self.classical, self.zpmotion = classical, zpmotion
        self.T, self.debye, self.cats = T, debye, N.array(cats, dtype='int')
        self.K00, self.K01, self.V01 = K00, K01, V01
        self.dt, self.nmd, self.ml = dt, nmd, ml
        self.kernel = None
        self.cids = N.array([[3 * c + 0, 3 * c + 1, 3 * c + 2] for c in cats], dtype='int').flatten()
        self.nc = len(self.cids)
        self.wmax = mcof * debye
        self.local = False
        self.nw = nw
        self.wl = [self.wmax * i / nw for i in range(nw)]          # phbath.py:57
        self.gamma, self.sig, self.gwl = gamma, sig, gwl
        self.eta_ad = eta_ad
        self._init_ensemble(batch, seed, device)
        if self.UseK():                                            # phbath.py:66-70
            raise NotImplementedError("phbath: Calculating self-energy is not implemented yet.")
        elif self.UseG() or self.UsePi():                          # phbath.py:72-81
            if self.UsePi():
                if len(self.sig[0]) != self.nc:
                    raise ValueError("phbath: inconsist between cids and sig!")
                self.ggamma()
            if self.UseG() and len(self.gamma[0]) != self.nc:
                raise ValueError("phbath: inconsist between cids and gamma!")
        else:                                                      # Debye model, phbath.py:82-89
            phfric = debye * N.pi / 6.0
            self.gamma = N.array([N.diag(phfric + N.zeros(int(self.nc)))])
            self.gwl = N.array([0])
            self.local = True
            self.ml = 1

    # -- setters of the reference -----------------------------------------------
    def SetMDsteps(self, dt, nmd):
        self.dt, self.nmd = dt, nmd
        self._noise_dev = None
        self._cur_dev = None

    def SetMemlen(self, len):
        self.ml = len

    def SetT(self, T):
        self.T = T

    def UseG(self):
        return self.gamma is not None and self.gwl is not None

    def UsePi(self):
        return self.sig is not None and self.gwl is not None

    def UseK(self):
        return self.K00 is not None and self.K01 is not None and self.V01 is not None

    # -- setup tables -------------------------------------------------------------
    def ggamma(self):
        """gamma(w) = -Im Sig(w)/w; the w == 0 entry is taken from the next grid point
        (phbath.py:126-144)."""
        if self.sig is None:
            raise ValueError("phbath.Gamma: self.sig is not set, need it to calculate gamma")
        sig, wl = N.asarray(self.sig), N.asarray(self.gwl, dtype=float)
        src = N.where(wl == 0, N.arange(len(wl)) + 1, N.arange(len(wl)))
        self.gamma = -N.imag(sig[src]) / wl[src][:, None, None]

    def gmem(self, device=None):
        """Friction kernel in time, K[k] = Re gamt(k dt) (phbath.py:161-190).  ``device``: run the
        cosine transform in libpymd_b200 (None: only when PYMD_SETUP_DEVICE=1)."""
        if self.ml is None or self.dt is None:
            raise ValueError("phbath.gmem: length of memory kernel not set!")
        if self.local:
            self.ml = 1
            self.kernel = N.asarray(self.gamma)
            return
        if device is None:
            device = self.setup_device if self.setup_device is not None else _noise.setup_on_device()
        tl = N.array([self.dt * i for i in range(self.ml)])
        gw = N.asarray(self.gwl, dtype=float)
        if device:
            self.kernel = gamt_device(tl, self.wl, self.gwl, self.gamma, self.eta_ad, self._dev())
        else:
            self.kernel = N.real(gamt(tl, self.wl, self.gwl, self.gamma, self.eta_ad))
        if self.eta_ad != 0:
            # gamma re-derived from the damped kernel: dt sum_t K(t) cos(w t)   (phbath.py:175-189)
            self.gammaOld = self.gamma
            if device:
                self.gamma = cosine_transform_device(gw, tl, self.kernel, 0.0, 0.5 * self.dt, self._dev())
            else:
                cosm = N.cos(gw[:, None] * tl[None, :])
                self.gamma = N.real(self.dt * N.tensordot(cosm, self.kernel, axes=(1, 0)))
            self._factors = None

    def _factor_key(self):
        return (self.T, self.wmax, self.dt, self.nmd, self.classical, self.zpmotion, id(self.gamma))

    def _spectral_factors(self):
        dev, strict = self._setup_target()
        return _noise.phonon_factors(self.gamma, self.gwl, self.T, self.wmax, self.dt, self.nmd,
                                     self.classical, self.zpmotion, device=dev, strict=strict,
                                     keep_vectors=self.setup_keep_vectors)

    def effective_selfenergy(self, w):
        """Retarded self-energy the time-domain friction actually implements, reconstructed from the
        discrete, truncated kernel: Sigma(w) = -i w dt sum_k K[k] exp(i w k dt) (no dt for the
        memory-less Debye bath) -- what harmonic.Drs must be fed for a tight comparison with MD spectra
        (SURVEY.md section 8c).  (len(w), nc, nc) complex."""
        w = N.asarray(w, dtype=float)
        ker = N.asarray(self.kernel, dtype=float)[:self.ml]
        if self.ml == 1:
            return -1j * w[:, None, None] * ker[0][None]
        ph = N.exp(1j * w[:, None] * (self.dt * N.arange(self.ml))[None, :])
        return -1j * (w * self.dt)[:, None, None] * N.tensordot(ph, ker, axes=(1, 0))

    def noise_spectrum(self, w):
        """Spectrum of the generated noise on the grid w: equ(w) gamma~(w) (noise.py:16-34, 61-66)."""
        w = N.asarray(w, dtype=float)
        return _noise.phnoisew(flinterp(w, self.gwl, N.asarray(self.gamma)), w, self.T, self.wmax,
                               self.classical, self.zpmotion)

    def gnoi(self, xi=None):
        """Regenerate the coloured noise for every trajectory (phbath.py:146-158); ``xi``
        optionally supplies the standard normals, shape (nmd/2+1, nc[, batch])."""
        self._gnoi(xi)

    def bforce(self, t, phis, qhis):
        """noise[t] - sum_k K[k] p_c(t-k) (dt) for one trajectory (phbath.py:192-202); inside
        md.vv the same contraction runs batched in libpymd_b200."""
        return self._bforce_host(t, phis, qhis)


def mf(f, cats, lens):
    """Scatter f into a zero vector of length lens (phbath.py:207-214)."""
    t = N.zeros(lens)
    t[N.asarray(cats, dtype=int)] = f
    return t


def exlist(a, indices):
    """Gather (phbath.py:217-218)."""
    return N.asarray(a)[N.asarray(indices, dtype=int)]


def gamt(tl, wl, gwl, gam, eta_ad=0):
    """gam(t) = 2 mean_i[gamma~(w_i) cos(w_i t)] w_max'/pi on the uniform grid wl, gamma~ by
    the reference's flinterp (phbath.py:221-254); with eta_ad the damped complex form."""
    tl = N.asarray(tl, dtype=float)
    w = N.asarray(wl, dtype=float)
    g = flinterp(w, gwl, N.asarray(gam))                       # (nw, nc, nc)
    nw = len(w)
    if eta_ad == 0:
        c = N.cos(tl[:, None] * w[None, :])                    # (ml, nw)
        return N.real(2.0 * N.tensordot(c, g, axes=(1, 0)) / nw * w[-1] / N.pi)
    with N.errstate(invalid="ignore", divide="ignore"):
        fm = w / (w - 1j * eta_ad)
        fp = w / (w + 1j * eta_ad)
    ph = tl[:, None] * w[None, :]
    dec = N.exp(-eta_ad * tl)[:, None]
    c = fm[None, :] * N.exp(-1j * ph) * dec + fp[None, :] * N.exp(+1j * ph) * dec
    return N.real(N.tensordot(c, g, axes=(1, 0)) / nw * w[-1] / N.pi)


def cosine_transform_device(tl, w, g, eta_ad, scale, device):
    """out[k] = scale sum_i coef(tl[k], w[i]) g[i] on the GPU (pymd_gamt); g is (len(w), ...) real,
    coef = 2 cos(t w) or its damped form.  Returns a host array of shape (len(tl),) + g.shape[1:]."""
    import torch
    g = N.ascontiguousarray(N.real(g), dtype=float)
    ne = int(N.prod(g.shape[1:])) if g.ndim > 1 else 1
    tl_d = torch.as_tensor(N.ascontiguousarray(tl, dtype=float), device=device)
    w_d = torch.as_tensor(N.ascontiguousarray(w, dtype=float), device=device)
    g_d = torch.as_tensor(g.reshape(len(w), ne), device=device)
    out = torch.empty((len(tl), ne), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        _lib.call("pymd_gamt", _lib.ptr(tl_d), len(tl), _lib.ptr(w_d), len(w), _lib.ptr(g_d), ne,
                  float(eta_ad), float(scale), _lib.ptr(out), _lib.stream_ptr())
    return out.cpu().numpy().reshape((len(tl),) + g.shape[1:])


def gamt_device(tl, wl, gwl, gam, eta_ad, device):
    """gamt() with the [ml x nw] x [nw x nc^2] product on the GPU; the interpolated table
    gamma~(w_i) is still the reference's flinterp, evaluated on the host."""
    w = N.asarray(wl, dtype=float)
    g = flinterp(w, gwl, N.asarray(gam))
    return cosine_transform_device(tl, w, g, eta_ad, w[-1] / N.pi / len(w), device)
