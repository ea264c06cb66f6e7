"""Electron bath with the reference's class interface (ebath.py:10-199): time-local friction
(ml = 1) plus the bias-dependent non-conservative, renormalisation and Berry forces."""
import numpy as N

from . import noise as _noise
from ._bathbase import BathBase
from .matrix import antisymmetrize, chkShape, symmetrize


class ebath(BathBase):
    """ebath(cats, T, dt, nmd, wmax=None, nw=None, bias=0., efric=None, exim=None, exip=None,
    zeta1=None, zeta2=None, classical=False, zpmotion=True)  -- ebath.py:30-31."""

    def __init__(self, cats, T, dt, nmd, wmax=None, nw=None, bias=0., efric=None, exim=None, exip=None,
                 zeta1=None, zeta2=None, classical=False, zpmotion=True, batch=None, seed=None, device=None):
        self.cats = N.array(cats, dtype='int')
        self.cids = N.array([[3 * c + 0, 3 * c + 1, 3 * c + 2] for c in cats], dtype='int').flatten()
        self.nc = len(self.cids)
        self.T, self.wmax = T, wmax
        self.nw, self.bias = nw, bias
        self.dt, self.nmd = dt, nmd
        self.classical, self.zpmotion = classical, zpmotion
        self.wl = None if (nw is None or wmax is None) else [self.wmax * i / nw for i in range(nw)]
        self._init_ensemble(batch, seed, device)
        self.CheckEmat(efric, exim, exip, zeta1, zeta2)
        self.ml = 1                                                # ebath.py:50
        self.local = True

    def CheckEmat(self, efric=None, exim=None, exip=None, zeta1=None, zeta2=None):
        """Shape checks; efric/exip/zeta1 symmetrised, exim/zeta2 antisymmetrised
        (ebath.py:56-130)."""
        if efric is None:
            self.efric = self.kernel = self.exim = self.exip = self.zeta1 = self.zeta2 = None
            self.ebath = False
            return
        n = chkShape(efric)
        if n != self.nc:
            raise ValueError("ebath.CheckEmat: efric shape error!")
        self.efric = symmetrize(efric)
        self.kernel = N.array([self.efric])
        self.exip, self.exim = N.zeros((n, n)), N.zeros((n, n))
        self.zeta1, self.zeta2 = N.zeros((n, n)), N.zeros((n, n))
        self.ebath = True
        for name, m, op in (("exim", exim, antisymmetrize), ("exip", exip, symmetrize),
                            ("zeta1", zeta1, symmetrize), ("zeta2", zeta2, antisymmetrize)):
            if m is not None:
                if chkShape(m) != self.nc:
                    raise ValueError("ebath.CheckEmat: the dimension of %s is wrong!" % name)
                setattr(self, name, op(m))
        self._factors = None

    def _factor_key(self):
        return (self.T, self.wmax, self.dt, self.nmd, self.bias, self.classical, self.zpmotion, id(self.efric))

    def _spectral_factors(self):
        dev, strict = self._setup_target()
        return _noise.electron_factors(self.efric, self.exim, self.exip, self.bias, self.T, self.wmax,
                                       self.dt, self.nmd, self.classical, self.zpmotion, device=dev, strict=strict,
                                       keep_vectors=self.setup_keep_vectors)

    def gnoi(self, xi=None):
        """Electronic noise for every trajectory (ebath.py:132-148)."""
        if self.ebath is False:
            raise ValueError("ebath.gnoi: ebath is False!")
        self._gnoi(xi)

    def GetSig(self):
        """Wide-band effective retarded self-energy (ebath.py:150-165)."""
        if self.wl is None:
            raise ValueError("ebath.GetSig:wl is not set")
        w = N.asarray(self.wl)[:, None, None]
        self.sig = -1.j * w * (self.efric + self.bias * self.zeta2) + self.bias * self.zeta1 - self.bias * self.exim

    def SetMDsteps(self, dt, nmd):
        self.dt, self.nmd = dt, nmd
        self._noise_dev = None
        self._cur_dev = None

    def setbias(self, bias=0.0):
        """BIAS IS mu_L - mu_R (ebath.py:172-178); the noise must be regenerated."""
        self.bias = bias

    # operands of ebath.bforce (ebath.py:191-194) pre-combined for the device:
    #   f = noise - (efric + V zeta2).p_c + V (exim - zeta1).q_c
    def effective_selfenergy(self, w):
        """-i w (efric + V zeta2) + V zeta1 - V exim (harmonic.SigE, harmonic.py:86-96)."""
        from .harmonic import SigE
        return SigE(w, self.efric, self.exim, self.zeta1, self.zeta2, self.bias)

    def noise_spectrum(self, w):
        """Spectrum of the generated noise on the grid w (noise.enoisew, noise.py:91-129)."""
        return _noise.enoisew(w, self.efric, self.exim, self.exip, self.bias, self.T, self.wmax,
                              self.classical, self.zpmotion)

    def _ap_matrix(self):
        return self.efric + self.bias * self.zeta2

    def _aq_matrix(self):
        if self.bias == 0 or (not N.any(self.exim) and not N.any(self.zeta1)):
            return None
        return self.bias * self.exim - self.bias * self.zeta1

    def bforce(self, t, phis, qhis):
        """ebath.py:180-199 for one trajectory given host history arrays."""
        import torch
        phis = N.asarray(phis, dtype=float)
        qhis = N.asarray(qhis, dtype=float)
        dev = self._dev()
        f = self._noise_dev[int(t) % self.nmd, :, 0].clone()
        f -= torch.as_tensor(self._ap_matrix(), dtype=torch.float64, device=dev) @ \
            torch.as_tensor(phis[0, self.cids], dtype=torch.float64, device=dev)
        aq = self._aq_matrix()
        if aq is not None:
            f += torch.as_tensor(aq, dtype=torch.float64, device=dev) @ \
                torch.as_tensor(qhis[0, self.cids], dtype=torch.float64, device=dev)
        out = N.zeros(phis.shape[1])
        out[self.cids] = f.cpu().numpy()
        return out


def mf(f, cats, lens):
    t = N.zeros(lens)
    t[N.asarray(cats, dtype=int)] = f
    return t


def exlist(a, indices):
    return N.asarray(a)[N.asarray(indices, dtype=int)]
