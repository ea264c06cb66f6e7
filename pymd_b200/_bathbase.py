"""Plumbing shared by phbath and ebath: ensemble attachment, device noise/current buffers.

The reference's bath protocol (what md touches): .dt, .nmd, .ml, .nc,
.noise, .cur, .gnoi(), .bforce(t, phis, qhis).  The extra optional knobs here (batch, seed,
xi=) default to the reference's single-trajectory behaviour.
"""
import numpy as N

from . import _lib
from . import noise as _noise


def _round_up(x, m):
    return (x + m - 1) // m * m


class BathBase(object):
    def _init_ensemble(self, batch=None, seed=None, device=None):
        self._index = 0                # position in md.baths (selects the Philox stream)
        self._batch = batch            # None until attached (md.AddBath) or first gnoi()
        self._seed = seed
        self._device = device
        self._traj0 = 0
        self._ngnoi = 0
        self._noise_dev = None
        self._cur_dev = None
        self._factors = None
        self._factors_key = None

    # -- ensemble bookkeeping -------------------------------------------------
    @property
    def batch(self):
        return 1 if self._batch is None else self._batch

    @property
    def ldb(self):
        return _round_up(self.batch, 32)

    def _attach(self, batch, seed, traj0, device, index=0):
        self._index = index
        if self._batch is not None and self._batch != batch:
            raise ValueError("bath was set up for an ensemble of %d, md has %d" % (self._batch, batch))
        self._batch = batch
        if self._seed is None:
            self._seed = seed
        self._traj0 = traj0
        self._device = device
        if self._noise_dev is not None and self._noise_dev.shape[2] != self.ldb:
            self._noise_dev = None     # generated before the ensemble size was known

    def _dev(self):
        import torch
        if self._device is None:
            self._device = torch.device("cuda", torch.cuda.current_device())
        return self._device

    # -- noise ------------------------------------------------------------------
    setup_device = None        # True: decompose the noise covariance on the GPU; None: PYMD_SETUP_DEVICE decides
    setup_keep_vectors = False  # also keep the eigenvectors of an on-device decomposition (spectral_factors().U)

    def _setup_target(self):
        """(device, strict) for the set-up tables: (None, ...) keeps the host numpy/LAPACK path, whose
        tables are bit-identical to the reference's."""
        if self.setup_device is None:
            return (self._dev() if _noise.setup_on_device() else None), False
        return (self._dev() if self.setup_device else None), True

    def _spectral_factors(self):
        raise NotImplementedError

    def spectral_factors(self):
        """Cached (lam, U, L) of the noise covariance; recomputed when T/bias/dt/nmd change."""
        dev, _ = self._setup_target()
        key = (self._factor_key(), None if dev is None else str(dev), self.setup_keep_vectors)
        if self._factors is None or self._factors_key != key:
            self._factors = self._spectral_factors()
            self._factors_key = key
        return self._factors

    def _gnoi(self, xi=None, segment=None, out=None, traj0=None):
        """``segment``: index of the md.Run segment the noise is for; selects the Philox stream, so that the
        realisation depends only on (seed, bath, segment, trajectory) and a resumed run repeats the uninterrupted
        one.  Hand-driven loops (segment=None) count the calls instead.  ``out`` / ``traj0``: generate into that
        device buffer for the sub-ensemble starting at that global trajectory index WITHOUT making it the current
        noise (ensemble.run_pipelined prepares the next sub-ensemble on a side stream this way)."""
        import torch
        if self.dt is None or self.nmd is None:
            raise ValueError("%s.gnoi: the md information dt and nmd are not set!" % type(self).__name__)
        fac = self.spectral_factors()
        nf = int(self.nmd / 2) + 1
        dev = self._dev()
        detached = out is not None
        if out is None:
            out = self._noise_target()
        if xi is None:
            if self._seed is None:
                self._seed = int(N.random.randint(0, 2 ** 31 - 1))
            sid = self._ngnoi if segment is None else int(segment)
            _noise.generate_philox(fac, self.dt, self.nmd, self.ldb, self.batch, self._seed,
                                   ((self._index + 1) << 16) ^ (sid & 0xFFFF),
                                   self._traj0 if traj0 is None else int(traj0), out)
        else:
            x = _noise._as_device_xi(xi, nf, self.nc, self.batch, self.ldb, dev)
            _noise.generate(fac, x, self.dt, self.nmd, out=out)
        wrap = getattr(self, "_noise_wrap", None)
        if wrap is not None and out.data_ptr() == self._noise_dev.data_ptr():
            wrap.copy_(out[0])              # overlaid records: "row nmd" of the noise is a copy of row 0 (phbath.py:196)
        if detached:
            return
        self._ngnoi = self._ngnoi + 1 if segment is None else int(segment) + 1
        self._noise_dev = out
        self._noise_version = getattr(self, "_noise_version", 0) + 1

    def _noise_target(self):
        """The device buffer the next gnoi() writes: the current one, or the spare one of a double-buffered run
        (ensemble.run_pipelined generates the next sub-ensemble's noise while the current one is stepping)."""
        import torch
        shape = (self.nmd, self.nc, self.ldb)
        if getattr(self, "_noise_wrap", None) is not None:
            return self._noise_dev          # a view into md's overlaid record buffer
        spare = getattr(self, "_noise_spare", None)
        if spare is not None and tuple(spare.shape) == shape:
            self._noise_spare = self._noise_dev if (self._noise_dev is not None and tuple(self._noise_dev.shape) == shape) else None
            return spare
        if self._noise_dev is None or tuple(self._noise_dev.shape) != shape:
            return torch.empty(shape, dtype=torch.float64, device=self._dev())
        return self._noise_dev

    @property
    def noise(self):
        """Host copy shaped like the reference's (nmd, nc); (nmd, nc, batch) for an ensemble."""
        if self._noise_dev is None:
            return None
        a = self._noise_dev[:, :, :self.batch].cpu().numpy()
        return a[:, :, 0] if self._batch in (None, 1) else a

    @noise.setter
    def noise(self, value):
        import torch
        if value is None:
            self._noise_dev = None
            return
        a = torch.as_tensor(N.asarray(value), dtype=torch.float64)
        if a.dim() == 2:
            a = a[:, :, None]
        if a.shape[0] != self.nmd or a.shape[1] != self.nc or a.shape[2] != self.batch:
            raise ValueError("noise must have shape (nmd, nc[, batch])")
        full = torch.zeros((self.nmd, self.nc, self.ldb), dtype=torch.float64, device=self._dev())
        full[:, :, :self.batch] = a.to(self._dev())
        self._noise_dev = full
        self._noise_version = getattr(self, "_noise_version", 0) + 1

    # -- heat current (written by md.vv, md.py:342) --------------------------
    def _alloc_cur(self):
        import torch
        if self._cur_dev is None or tuple(self._cur_dev.shape) != (self.nmd, self.ldb):
            self._cur_dev = torch.zeros((self.nmd, self.ldb), dtype=torch.float64, device=self._dev())
        return self._cur_dev

    @property
    def cur(self):
        if self._cur_dev is None:
            return N.zeros(self.nmd)
        a = self._cur_dev[:, :self.batch].cpu().numpy()
        return a[:, 0] if self._batch in (None, 1) else a

    # -- single-trajectory host-callable force (API parity with bath.bforce) -----
    def _bforce_host(self, t, phis, qhis):
        """noise[t] - friction for ONE trajectory given host history arrays; used only when a
        script calls bath.bforce by hand.  Runs the contraction on the GPU."""
        import torch
        phis = N.asarray(phis, dtype=float)
        qhis = N.asarray(qhis, dtype=float)
        dev = self._dev()
        nz = self._noise_dev[int(t) % self.nmd, :, 0]
        ker = torch.as_tensor(N.asarray(self.kernel), dtype=torch.float64, device=dev)
        ph = torch.as_tensor(phis[:, self.cids], dtype=torch.float64, device=dev)
        f = nz.clone()
        fr = torch.einsum("kij,kj->i", ker[:self.ml], ph[:self.ml])
        f -= fr * (self.dt if self.ml > 1 else 1.0)
        aq = self._aq_matrix()
        if aq is not None:
            f += torch.as_tensor(aq, dtype=torch.float64, device=dev) @ \
                torch.as_tensor(qhis[0, self.cids], dtype=torch.float64, device=dev)
        out = N.zeros(phis.shape[1])
        out[self.cids] = f.cpu().numpy()
        return out

    def _aq_matrix(self):
        return None
