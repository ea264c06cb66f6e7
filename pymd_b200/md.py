"""Langevin molecular dynamics with the reference's ``md`` class interface (md.py:23-705),
stepping an ENSEMBLE of independent trajectories on one B200 through libpymd_b200.

Same constructor, methods and attributes as the reference, so driver scripts written for
pymd (tests/test.py style hand loops, examples/*.py ending in ``mdrun.Run()``) drive it
unchanged.  New optional keyword arguments, all defaulting to the reference behaviour of
one trajectory: ``batch`` (ensemble size), ``seed``, ``device``, ``traj0`` (global index of
the first local trajectory, for multi-GPU sharding), ``step_mode``, ``record_fhis``.

There is no CPU fallback: the time loop, noise generation and spectra run in CUDA.
"""
import ctypes as C
import os
import time

import numpy as N
from numpy import linalg as LA

from . import _lib
from . import units as U
from .functions import bose
from .matrix import chkShape, symmetrize


def _round_up(x, m):
    return (x + m - 1) // m * m


class md(object):
    def __init__(self, dt, nmd, T, syslist=None, axyz=None, harmonic=False, dyn=None, savepq=True,
                 nrep=1, npie=8, constr=None, batch=1, seed=None, device=None, traj0=0, step_mode=0,
                 record_fhis=False):
        self.sint = None
        self.brennerrun = None
        self.lammpsrun = None
        self.constraint = constr
        self.nrep = nrep
        self.dt, self.nmd = dt, nmd
        self.harmonic = harmonic
        self.T = T
        self.npie = npie
        self.power = N.zeros((self.nmd, 2))
        self.power2 = N.zeros((self.nmd, 2))
        self.batch = int(batch)
        self.ldb = _round_up(self.batch, 32)
        self.seed = seed
        self.traj0 = int(traj0)
        self.step_mode = step_mode
        self.record_fhis = record_fhis
        self.write_files = True
        self.outdir = "."
        self.harmonic_check = False     # Run() also writes harmonic.<T>.run<j>.dat: w, analytic PP, MD power2
        self._device = device
        self.SetXyz(axyz)
        self.syslist = self.na = self.nph = None
        if syslist is not None:                                     # md.py:79-89
            if len(syslist) > self.nta or min(syslist) < 0 or max(syslist) > self.nta - 1:
                raise ValueError("syslist out of range")
            self.syslist = N.array(syslist, dtype='int')
            self.na = len(syslist)
            self.nph = 3 * len(syslist)
        elif axyz is not None:                                      # md.py:90-95
            self.syslist = N.array(range(len(axyz)), dtype='int')
            self.na = len(self.syslist)
            self.nph = 3 * len(self.syslist)
        self.ml = 1
        self.t = 0
        self.baths = []
        self._ctx = None
        self._ctx_key = None
        self._p = self._q = None
        self._ps = self._qs = self._etot = None
        self._rings = {}
        self._fhis = {}
        self._ninit = 0
        self.setDyn(dyn)
        self.savepq = savepq

    # ------------------------------------------------------------------ plain setters
    def info(self):
        print("--------------------------------------------")
        print("Basis information of the MD simulation:")
        print("Harmonic force: %s" % self.harmonic)
        print("System atom number: %s" % self.na)
        print("MD time step: %s" % self.dt)
        print("MD number of steps: %s" % self.nmd)
        print("MD memory kernel length: %s" % self.ml)
        print("Number of baths attached: %s" % len(self.baths))
        print("Ensemble size (trajectories on this GPU): %s" % self.batch)
        if self.dyn is None:
            print("md.info: No dynamical matrix input!")

    def SetT(self, T):
        self.T = T

    def SetMD(self, dt, nmd):
        self.dt, self.nmd = dt, nmd
        self._drop_ctx()

    def SetHarm(self, harmonic):
        self.harmonic = harmonic

    def SetXyz(self, axyz):
        if axyz is not None:
            self.xyz = N.array([a[1:] for a in axyz], dtype='d').flatten()
            self.els = [a[0] for a in axyz]
            self.nta = len(axyz)
        else:
            self.xyz = self.els = self.nta = None

    def SetSyslist(self, syslist):
        self.syslist = N.array(syslist)
        self.na = len(syslist)
        self.nph = 3 * len(syslist)
        if self.xyz is not None and len(self.syslist) > self.nta:
            raise ValueError("md.SetSyslist:system atom number larger than total atom number")

    def setDyn(self, dyn=None):
        """Symmetrise, clamp negative eigenvalues to zero and rebuild the dynamical matrix
        from its eigen-decomposition (md.py:211-250).  Host numpy, as in the reference."""
        if dyn is None:
            self.dyn = self.hw = self.U = None
            return
        ndyn = N.array(dyn)
        n = chkShape(ndyn)
        if self.nph is not None and self.nph != n:
            raise ValueError("md.setDyn: the dimension of dynamical matrix is wrong!")
        self.nph = n
        av, au = LA.eigh(symmetrize(ndyn))
        av = N.where(av < 0, 0.0, av)
        self.hw = N.sqrt(av)
        self.U = N.array(au)
        self.dyn = (self.U * av) @ self.U.T
        self._drop_ctx()

    def AddBath(self, bath):
        """md.py:160-176."""
        if self.dt != bath.dt:
            raise ValueError("md.AddBath: md time step dt not consistent!")
        if self.nmd != bath.nmd:
            raise ValueError("md.AddBath: number of md steps nmd not consistent!")
        bath._attach(self.batch, None if self.seed is None else self.seed + 7919 * (len(self.baths) + 1),
                     self.traj0, self._dev(), index=len(self.baths))
        self.baths.append(bath)
        if bath.ml > self.ml:
            self.ml = bath.ml
        self._drop_ctx()

    def AddSint(self, sint):
        self.sint = sint

    def AddBint(self, bint):
        self.brennerrun = bint

    def AddLMPint(self, lint):
        self.lammpsrun = lint

    # ------------------------------------------------------------------ device plumbing
    def _dev(self):
        import torch
        if self._device is None:
            self._device = torch.device("cuda", torch.cuda.current_device())
        elif not isinstance(self._device, torch.device):
            self._device = torch.device(self._device)
        return self._device

    def _drop_ctx(self):
        if getattr(self, "_ctx", None) is not None:
            _lib.load().pymd_ctx_destroy(self._ctx)
        self._ctx = None

    def __del__(self):
        try:
            self._drop_ctx()
        except Exception:
            pass

    def _zeros(self, *shape):
        import torch
        return torch.zeros(shape, dtype=torch.float64, device=self._dev())

    def _ensure(self):
        """Create the library context and device arrays once the configuration is known."""
        if self.sint is not None or self.brennerrun is not None or self.lammpsrun is not None:
            raise NotImplementedError(
                "external force drivers (SIESTA/Brenner/LAMMPS) need a per-step host callback and are "
                "outside the GPU stepping path; attach a dynamical matrix instead")
        if self.nph is None:
            raise ValueError("md: number of degrees of freedom unknown (give axyz, syslist or dyn)")
        key = self._operand_key()
        saved = None
        if self._ctx is not None:
            if key == self._ctx_key:
                return
            # an operand changed after the context was built (ebath.setbias, a re-run gmem()/SetMemlen, a new
            # constraint): the reference reads bias, kernel and constraint live on every force call
            # (ebath.py:191-194, phbath.py:197-201, md.py:355-358), so rebuild the device operands and keep
            # the velocity history
            saved = self._history_dev()
            self._drop_ctx()
        self._ctx_key = key
        import torch
        lib = _lib.load()
        dev = self._dev()
        torch.cuda.set_device(dev)
        nd, ldb = self.nph, self.ldb
        ctx = C.c_void_p()
        _lib.call("pymd_ctx_create", nd, ldb, len(self.baths), float(self.dt), int(self.nmd),
                  dev.index or 0, C.byref(ctx))
        self._ctx = ctx
        if self.dyn is not None:
            d = N.ascontiguousarray(self.dyn, dtype=N.float64)
            _lib.call("pymd_set_dyn", ctx, d.ctypes.data)
        if self.constraint is not None:
            c = N.ascontiguousarray(N.array(self.constraint, dtype=N.float64))
            if c.ndim != 2 or c.shape[1] != nd:
                raise ValueError("ApplyConstraint:shape error")
            _lib.call("pymd_set_constraints", ctx, c.ctypes.data, c.shape[0])
        if self._p is None or tuple(self._p.shape) != (nd, ldb):
            self._p, self._q = self._zeros(nd, ldb), self._zeros(nd, ldb)
        _lib.call("pymd_bind_state", ctx, self._p.data_ptr(), self._q.data_ptr())
        if self._etot is None or tuple(self._etot.shape) != (self.nmd, ldb):
            self._etot = self._zeros(self.nmd, ldb)
        for b, bath in enumerate(self.baths):
            if bath.kernel is None:
                raise ValueError("md: bath %d has no memory kernel (call bath.gmem() before AddBath)" % b)
            ker = N.ascontiguousarray(N.asarray(bath.kernel, dtype=N.float64)[:bath.ml])
            cids = N.ascontiguousarray(bath.cids, dtype=N.int32)
            aq = bath._aq_matrix()
            if hasattr(bath, "_ap_matrix"):
                ker = N.ascontiguousarray(bath._ap_matrix()[None, :, :], dtype=N.float64)
            aqc = None if aq is None else N.ascontiguousarray(aq, dtype=N.float64)
            _lib.call("pymd_bath_set", ctx, b, bath.nc, cids.ctypes.data, int(bath.ml), ker.ctypes.data,
                      None if aqc is None else aqc.ctypes.data, 1 if bath.ml > 1 else 0)
            if bath.ml > 1:
                R = lib.pymd_ring_len(int(bath.ml), 0)
                ncp = lib.pymd_bath_ncp(bath.nc)
                if b not in self._rings or tuple(self._rings[b].shape) != (R, ncp, ldb):
                    self._rings[b] = self._zeros(R, ncp, ldb)
                _lib.call("pymd_bind_history", ctx, b, self._rings[b].data_ptr(), R)
            if self.record_fhis:
                self._fhis[b] = self._zeros(self.nmd, bath.nc, ldb)
        self._savepq_alloc()
        if saved:
            for b, buf in saved.items():
                bath = self.baths[b]
                if bath.ml > 1 and tuple(buf.shape) == (bath.ml, bath.nc, ldb):
                    _lib.call("pymd_history_set", ctx, b, buf.data_ptr(), int(bath.ml), _lib.stream_ptr())

    def _operand_key(self):
        """Everything pymd_bath_set / pymd_set_constraints copied into the context.  Arrays are identified by
        object (gmem() and CheckEmat() make new ones); the small constraint matrix by value."""
        con = None if self.constraint is None else N.asarray(self.constraint, dtype=N.float64).tobytes()
        baths = tuple((id(b), b.nc, int(b.ml), id(b.kernel), getattr(b, "bias", None), id(getattr(b, "efric", None)),
                       id(getattr(b, "exim", None)), id(getattr(b, "zeta1", None)), id(getattr(b, "zeta2", None)))
                      for b in self.baths)
        return (self.nph, self.ldb, float(self.dt), int(self.nmd), con, baths, bool(self.record_fhis))

    def _history_dev(self):
        """Device copies [ml][nc][ldb] (newest first) of every non-local bath's history, keyed by bath."""
        out = {}
        for b, bath in enumerate(self.baths):
            if bath.ml > 1 and b in self._rings:
                try:
                    buf = self._zeros(bath.ml, bath.nc, self.ldb)
                    _lib.call("pymd_history_get", self._ctx, b, buf.data_ptr(), int(bath.ml), _lib.stream_ptr())
                    out[b] = buf
                except _lib.PymdError:
                    pass            # the ring no longer matches (ml grew): the history starts again from zero
        return out

    def _savepq_alloc(self):
        if getattr(self, "_overlay", False):
            return
        if self._savepq and self.nph is not None:
            shape = (self.nmd, self.nph, self.ldb)
            if self._ps is None or tuple(self._ps.shape) != shape:
                self._ps, self._qs = self._zeros(*shape), self._zeros(*shape)
        else:
            self._ps = self._qs = None

    # ------------------------------------------------------------------ overlaid records (large ensembles)
    RECORD_SLACK_ROWS = 36        # a launch covers <= 32 steps and its CTAs advance independently

    def enable_record_overlay(self):
        """Keep the saved trajectories AND the noise of every bath in ONE device buffer: row t of [ps | qs] is written
        exactly when noise row t has been consumed, so the 2 nph ldb doubles of a record row may overlay the
        end-aligned noise rows (sum_b nc_b ldb <= 2 nph ldb doubles each).  4096 trajectories of config 2 then need
        129 GB instead of 218 GB and step as ONE ensemble with savepq=True (the reference keeps both arrays for its
        single trajectory, md.py:146-147, phbath.py:157).  Consequences: bath.noise is only valid until the next
        steps() call; a steps() call must end at the end of the noise table (md.Run's pieces do); ResetSavepq does
        not zero; white noise cannot be injected (gnoi(xi=...)); the two-unit step kernel must apply (nph <= 64, small
        baths of <= 16 dofs).  Call after the last AddBath and before gnoi()."""
        import torch
        if not self._savepq:
            raise ValueError("md.enable_record_overlay: needs savepq=True")
        self._ps = self._qs = None
        self._overlay = True                # from here on _savepq_alloc leaves the records alone
        self._ensure()
        nd, ldb, nmd = self.nph, self.ldb, self.nmd
        W = 2 * nd * ldb                                   # doubles per record row
        V = sum(b.nc for b in self.baths) * ldb            # doubles per noise row
        if V > W:
            self._overlay = False
            raise ValueError("md.enable_record_overlay: the baths have more noise components than 2 nph")
        rows = nmd + self.RECORD_SLACK_ROWS
        self._ps = self._qs = None
        for b in self.baths:
            b._noise_dev = None
        torch.cuda.empty_cache()
        buf = torch.empty(rows * W, dtype=torch.float64, device=self._dev())
        self._record_buf = buf
        self._ps = buf.as_strided((nmd, nd, ldb), (W, ldb, 1), 0)
        self._qs = buf.as_strided((nmd, nd, ldb), (W, ldb, 1), nd * ldb)
        base = rows * W - (nmd + 1) * V                    # noise rows 0 .. nmd (row nmd = copy of row 0), end-aligned
        off = 0
        for b in self.baths:
            b._noise_dev = buf.as_strided((nmd, b.nc, ldb), (V, ldb, 1), base + off)
            b._noise_wrap = buf.as_strided((b.nc, ldb), (ldb, 1), base + nmd * V + off)
            b._noise_version = getattr(b, "_noise_version", 0) + 1
            off += b.nc * ldb

    def _bind(self):
        """(Re)bind every borrowed device array; cheap, done before each library call."""
        ctx = self._ctx
        if getattr(self, "_overlay", False):
            _lib.call("pymd_bind_outputs_strided", ctx, self._ps.data_ptr(), self._qs.data_ptr(), self._etot.data_ptr(),
                      int(self._ps.stride(0)))
            _lib.call("pymd_set_noise_wrap_row", ctx, int(self.nmd))
            for b, bath in enumerate(self.baths):
                _lib.call("pymd_bind_noise_strided", ctx, b, bath._noise_dev.data_ptr(), int(bath._noise_dev.stride(0)))
                _lib.call("pymd_bind_bath_outputs", ctx, b, bath._alloc_cur().data_ptr(), _lib.ptr(self._fhis.get(b)))
            return
        _lib.call("pymd_bind_outputs", ctx, _lib.ptr(self._ps), _lib.ptr(self._qs), self._etot.data_ptr())
        for b, bath in enumerate(self.baths):
            if bath._noise_dev is None:
                raise ValueError("md.vv: bath %d has no noise, call bath.gnoi() first" % b)
            if tuple(bath._noise_dev.shape) != (self.nmd, bath.nc, self.ldb):
                raise ValueError("md.vv: bath %d noise has the wrong shape (ensemble %d)" % (b, self.batch))
            _lib.call("pymd_bind_noise", ctx, b, bath._noise_dev.data_ptr())
            _lib.call("pymd_bind_bath_outputs", ctx, b, bath._alloc_cur().data_ptr(), _lib.ptr(self._fhis.get(b)))

    # ------------------------------------------------------------------ state access
    @property
    def savepq(self):
        return self._savepq

    @savepq.setter
    def savepq(self, v):
        self._savepq = bool(v)
        if not v:
            self._ps = self._qs = None

    def ResetSavepq(self, savepq=True):
        """md.py:143-150: (re)allocate and zero the saved trajectories."""
        self._savepq = bool(savepq)
        if self.nmd is None or self.nph is None:
            raise ValueError("md.ResetSavepq: nmd or nph is not set!")
        if getattr(self, "_overlay", False):
            return                      # the record rows share memory with the noise: every row is rewritten anyway
        self._savepq_alloc()
        if self._ps is not None:
            self._ps.zero_()
            self._qs.zero_()

    def _host(self, t, lead=True):
        """device [..., ldb] -> host array in the reference's shape, with a trailing ensemble
        axis when batch > 1."""
        a = t[..., :self.batch].cpu().numpy()
        return a[..., 0] if self.batch == 1 else a

    def _upload(self, dst, value):
        """host array in the reference's shape (broadcast over the ensemble) or with a trailing
        ensemble axis -> device [..., ldb]."""
        import torch
        a = N.asarray(value, dtype=N.float64)
        want = tuple(dst.shape[:-1])
        if a.shape == want:
            a = N.broadcast_to(a[..., None], want + (self.batch,))
        elif a.shape != want + (self.batch,):
            raise ValueError("expected shape %s or %s, got %s" % (want, want + (self.batch,), a.shape))
        dst.zero_()
        dst[..., :self.batch] = torch.as_tensor(N.array(a, dtype=N.float64, order='C'), dtype=torch.float64).to(dst.device)

    @property
    def p(self):
        return [] if self._p is None else self._host(self._p, True)

    @p.setter
    def p(self, v):
        self._ensure()
        self._upload(self._p, v)
        _lib.call("pymd_invalidate_cache", self._ctx)

    @property
    def q(self):
        return [] if self._q is None else self._host(self._q, True)

    @q.setter
    def q(self, v):
        self._ensure()
        self._upload(self._q, v)
        _lib.call("pymd_invalidate_cache", self._ctx)

    @property
    def ps(self):
        return None if self._ps is None else self._host(self._ps, True)

    @property
    def qs(self):
        return None if self._qs is None else self._host(self._qs, True)

    @property
    def etot(self):
        return N.zeros(self.nmd) if self._etot is None else self._host(self._etot, True)

    @property
    def fhis(self):
        """Bath forces of the first evaluation of each step (md.py:343), restricted to the
        bath's dofs and only when ``record_fhis=True`` (the reference keeps nmd x nph per bath
        per trajectory, which cannot be materialised for an ensemble)."""
        out = []
        for b, bath in enumerate(self.baths):
            if b not in self._fhis:
                out.append(None)
                continue
            loc = self._host(self._fhis[b])
            full = N.zeros((self.nmd, self.nph) + loc.shape[2:])
            full[:, bath.cids] = loc
            out.append(full)
        return out

    @property
    def fbaths(self):
        """Bath forces at the current state (md.py:380 leaves the last evaluated ones in md.fbaths): the first
        force evaluation of the next step, for trajectory 0 of an ensemble.  Diagnostic, evaluated on demand."""
        if self._ctx is None or any(b._noise_dev is None for b in self.baths):
            return [None] * len(self.baths)
        phis = self._history() if self.batch == 1 else self._history()[..., 0]
        p = self.p if self.batch == 1 else self.p[..., 0]
        q = self.q if self.batch == 1 else self.q[..., 0]
        qhis = N.zeros_like(phis)
        phis[0], qhis[0] = p, q
        return [b.bforce(self.t, phis, qhis) for b in self.baths]

    @property
    def phis(self):
        """Velocity history, newest first, shape (ml, nph) like the reference (md.py:296-298);
        only the columns a non-local bath reads are kept on the device."""
        return self._history()

    def _history(self):
        self._ensure()
        out = N.zeros((self.ml, self.nph) + ((self.batch,) if self.batch > 1 else ()))
        for b, bath in enumerate(self.baths):
            if bath.ml <= 1:
                continue
            buf = self._zeros(bath.ml, bath.nc, self.ldb)
            _lib.call("pymd_history_get", self._ctx, b, buf.data_ptr(), int(bath.ml), _lib.stream_ptr())
            out[:bath.ml, bath.cids] = self._host(buf)
        return out

    def set_history(self, phis):
        """Restore the velocity history (resume path, md.py:550-551/576-582); phis has the
        reference's shape (ml, nph) [+ trailing ensemble axis]."""
        self._ensure()
        phis = N.asarray(phis, dtype=N.float64)
        for b, bath in enumerate(self.baths):
            if bath.ml <= 1:
                continue
            buf = self._zeros(bath.ml, bath.nc, self.ldb)
            self._upload(buf, phis[:bath.ml][:, bath.cids])
            _lib.call("pymd_history_set", self._ctx, b, buf.data_ptr(), int(bath.ml), _lib.stream_ptr())

    # ------------------------------------------------------------------ the reference's methods
    def energy(self):
        """Kinetic energy 0.5 p.p of the current state (md.py:154-158)."""
        import torch
        e = 0.5 * (self._p[:, :self.batch] ** 2).sum(0).cpu().numpy()
        return float(e[0]) if self.batch == 1 else e

    def initialise(self, r=None, stream=None):
        """Quantum initial displacements/velocities from the eigenmodes (md.py:253-290).
        ``r``: optional uniforms in [0,1), shape (nph,) or (nph, batch); by
        default one Philox stream per trajectory (``stream``: which one; default a per-object call counter)."""
        import torch
        self._ensure()
        self.t = 0
        if self.dyn is None:
            self._p.zero_()
            self._q.zero_()
            _lib.call("pymd_invalidate_cache", self._ctx)
            return
        nd = self.nph
        hw = N.ascontiguousarray(self.hw, dtype=N.float64)
        with N.errstate(divide="ignore", invalid="ignore"):
            am = N.where(hw < 0.01, 0.0, ((bose(hw, self.T) + 0.5) * 2.0 / N.where(hw > 0, hw, 1.0)) ** 0.5)
        am = N.ascontiguousarray(am, dtype=N.float64)
        rdev = self._zeros(nd, self.ldb)
        if r is None:
            seed = self.seed if self.seed is not None else int(N.random.randint(0, 2 ** 31 - 1))
            sid = self._ninit if stream is None else int(stream)
            _lib.call("pymd_philox_uniform", rdev.data_ptr(), nd, self.ldb, int(seed), 0xFFFF0000 + (sid & 0xFFFF),
                      self.traj0, _lib.stream_ptr())
            if stream is None:
                self._ninit += 1
        else:
            ra = N.asarray(r, dtype=N.float64)
            if ra.shape == (nd,):
                ra = N.broadcast_to(ra[:, None], (nd, self.batch))
            elif ra.shape != (nd, self.batch):
                raise ValueError("r must have shape (nph,) or (nph, batch)")
            rdev[:, :self.batch] = torch.as_tensor(N.ascontiguousarray(ra)).to(rdev.device)
        Um = N.ascontiguousarray(self.U, dtype=N.float64)
        _lib.call("pymd_init_state", self._ctx, Um.ctypes.data, hw.ctypes.data, am.ctypes.data,
                  rdev.data_ptr(), _lib.stream_ptr())
        if self.batch < self.ldb:
            self._p[:, self.batch:] = 0.0
            self._q[:, self.batch:] = 0.0
        self.pinit, self.qinit = self.p, self.q

    def ResetHis(self):
        """Zero the friction-kernel history (md.py:292-301)."""
        if self.nph is None or self.ml is None:
            raise ValueError("self.nph and self.ml are not set!")
        self._ensure()
        _lib.call("pymd_reset_history", self._ctx, _lib.stream_ptr())

    def steps(self, n):
        """n consecutive md.vv steps for the whole ensemble in one library call."""
        self._ensure()
        self._bind()
        _lib.call("pymd_step", self._ctx, int(self.t), int(n), int(self.step_mode), _lib.stream_ptr())
        self.t = int(self.t) + int(n)

    def vv(self):
        """One modified velocity-Verlet step (md.py:315-358)."""
        self.steps(1)

    def potforce(self, q):
        """-dyn.q for a host vector (md.py:444-492, dynamical-matrix branch)."""
        import torch
        if self.dyn is None:
            raise ValueError("no driver, no md")
        d = torch.as_tensor(self.dyn, dtype=torch.float64, device=self._dev())
        return (-(d @ torch.as_tensor(N.asarray(q, dtype=float), device=self._dev()))).cpu().numpy()

    def force(self, t, p, q, id=0):
        """Potential + bath forces for ONE host state (md.py:360-383), using the current device
        history; diagnostic helper (the time loop never calls it)."""
        phis = self._history() if self.batch == 1 else self._history()[..., 0]
        qhis = N.zeros_like(phis)
        if id == 0:
            phis[0], qhis[0] = p, q
        else:
            phis = N.concatenate((N.asarray([p]), phis[:-1]), axis=0)
            qhis[0] = q
        f = self.potforce(q)
        for b in self.baths:
            f = f + b.bforce(t + id, phis, qhis)
        return f

    def GetPower(self):
        """Ensemble-averaged power spectra from the saved trajectories (md.py:303-313,
        spectrum.py:8-45)."""
        if not self.savepq or self._ps is None:
            raise ValueError("md.GetPower: trajectories not saved! you need to set savepq to True!")
        from .spectrum import ensemble_powerspec
        dw = 2. * N.pi / self.dt / self.nmd
        w = dw * N.arange(self.nmd)
        s1 = ensemble_powerspec(self._qs, self.dt, self.nmd, self.batch, wsq=True)
        s2 = ensemble_powerspec(self._ps, self.dt, self.nmd, self.batch, wsq=False)
        self.power_sum, self.power2_sum = s1, s2              # sums over local trajectories (for allreduce)
        self.power = N.stack((w, s1 / self.batch), axis=1)
        self.power2 = N.stack((w, s2 / self.batch), axis=1)

    def HarmonicSpectra(self, wl=None):
        """Analytic counterpart of ``power2`` for the attached baths (the reference's harmonic.py
        cross-check, harmonic.py:37-84): D(w) = [(w + i eta)^2 - dyn - sum_b Sigma_b(w)]^-1 with every
        bath's self-energy as the stepping path implements it (bath.effective_selfenergy) and
        PP(w) = w^2 Tr[D Pi D^+] with Pi = sum_b noise spectrum; all frequencies in one GPU launch
        (pymd_harmonic).  ``wl`` defaults to the MD frequencies 0 < w < 1.5 max(hw).
        Returns dict(w, DDos, UU, PP); in the steady state power2[:, 1] ~ PP on the same grid."""
        from . import harmonic as H
        if self.dyn is None:
            raise ValueError("md.HarmonicSpectra: needs the dynamical matrix")
        if wl is None:
            w = 2. * N.pi / self.dt / self.nmd * N.arange(1, self.nmd // 2)
            wl = w[w < 1.5 * max(self.hw)]
        wl = N.asarray(wl, dtype=float)
        if len(wl) == 0:
            raise ValueError("md.HarmonicSpectra: no MD frequency below 1.5 max(hw); the run is too short")
        nd = self.nph
        se = N.zeros((len(wl), nd, nd), dtype=complex)
        pi = N.zeros((len(wl), nd, nd), dtype=complex)
        for b in self.baths:
            c = N.asarray(b.cids)
            se[:, c[:, None], c[None, :]] += b.effective_selfenergy(wl)
            pi[:, c[:, None], c[None, :]] += b.noise_spectrum(wl)
        ddos, uu, pp = H.spectra_device(wl, self.dyn, se, pi, self._dev())
        return dict(w=wl, DDos=ddos, UU=uu, PP=pp)

    def Run(self):
        """nrep consecutive segments of nmd steps each: fresh noise per segment, npie pieces
        with a checkpoint after each, running mean of the spectra, kappa/power text outputs
        (md.py:523-652)."""
        self.initialise(stream=0)
        self.ResetHis()
        self.curmean = N.zeros((self.nrep, len(self.baths)))
        for j in range(self.nrep):
            ipie = -1
            state = self._resume(j)
            if state == "finished":
                continue
            if state is not None:
                ipie = state
            else:
                for b in self.baths:
                    # noise realisation = f(seed, bath, segment): a resumed run draws what the uninterrupted one would
                    b._gnoi(segment=j) if hasattr(b, "_gnoi") else b.gnoi()
                self.ResetSavepq(self._savepq)
            for i in range(ipie + 1, self.npie):
                self.steps(self.nmd // self.npie)
                self.dump(i, j)
            power, power2 = N.copy(self.power), N.copy(self.power2)
            if self._savepq:
                self.GetPower()
                self.power = (power * j + self.power) / float(j + 1)       # md.py:604-607
                self.power2 = (power2 * j + self.power2) / float(j + 1)
            self.dump(self.npie - 1, j)
            from .spectrum import time_means
            for ii, b in enumerate(self.baths):
                self.curmean[j, ii] = float(time_means(b._cur_dev)[:self.batch].mean().item()) * U.curcof
            if self.write_files:
                self._write_outputs(j)

    # ------------------------------------------------------------------ outputs / checkpoints
    def _write_outputs(self, j):
        """kappa.<T>.bath<i>.run<j>.dat, power.<T>.run<j>.dat, power2.<T>.run<j>.dat
        (md.py:616-652); spectra truncated at 1.5 max(hw)."""
        for ii in range(len(self.baths)):
            with open(os.path.join(self.outdir, "kappa.%s.bath%d.run%d.dat" % (self.T, ii, j)), "w") as fk:
                fk.write("%i %f    %f \n" % (j, self.T, self.curmean[j, ii]))
        for name, arr in (("power", self.power), ("power2", self.power2)):
            with open(os.path.join(self.outdir, "%s.%s.run%d.dat" % (name, self.T, j)), "w") as f:
                lim = None if self.hw is None else 1.5 * max(self.hw)
                for i in range(len(arr)):
                    if lim is not None and not arr[i, 0] < lim:
                        break
                    f.write("%f     %f \n" % (arr[i, 0], arr[i, 1]))
        if self.harmonic_check and self._savepq:
            # the reference's cross-check (harmonic.py:70-84: MD power2 ~ PP) beside the spectra
            self.harmonic_spectra = h = self.HarmonicSpectra()
            k = N.rint(h["w"] / (2. * N.pi / self.dt / self.nmd)).astype(int)
            with open(os.path.join(self.outdir, "harmonic.%s.run%d.dat" % (self.T, j)), "w") as f:
                for i in range(len(k)):
                    f.write("%f     %e     %e \n" % (h["w"][i], h["PP"][i], self.power2[k[i], 1]))

    def _ckpt_name(self, j):
        """Single trajectories checkpoint into the reference's own container, ``MD<j>.nc`` (NetCDF-3 with the
        variables and dimensions of md.py:654-705, written with scipy); ensembles add a trailing axis that the
        reference's schema has no dimension for and go to ``MD<j>.npz``."""
        return os.path.join(self.outdir, ("MD%d.nc" if self.batch == 1 else "MD%d.npz") % j)

    # ensembles whose full checkpoint (noise, saved trajectories, currents) would exceed this many bytes write a
    # state-only checkpoint at the END of each segment instead of the reference's everything-after-every-piece
    checkpoint_bytes_limit = 2 << 30

    def _full_checkpoint_bytes(self):
        per_traj = sum(b.nc + 1 for b in self.baths) + 1 + (2 * self.nph if self._savepq else 0)
        if self.record_fhis:
            per_traj += sum(b.nc for b in self.baths)
        return 8 * self.nmd * per_traj * self.batch

    def dump(self, ipie, id):
        """Checkpoint with the reference's variable names (md.py:654-705).  The reference rewrites noise, saved
        trajectories and currents after every piece; for a large ensemble that is hundreds of GB per dump, so
        above ``checkpoint_bytes_limit`` only the end of a segment is checkpointed, and only what chaining the
        next segment and skipping a finished one need (p, q, t, ipie, phis, power, power2, per-trajectory mean
        currents and kinetic energy).  The noise of a segment is a pure function of (seed, bath, segment,
        trajectory), so nothing else has to be stored to continue the run."""
        if not self.write_files:
            return
        light = self.batch > 1 and self._full_checkpoint_bytes() > self.checkpoint_bytes_limit
        if light and ipie + 1 < self.npie:
            return
        d = dict(power=self.power, power2=self.power2, p=self.p, q=self.q,
                 t=N.array([self.t]), ipie=N.array([ipie]), phis=self._history())
        if light:
            d["light"] = N.array([1])
            d["energy_mean"] = self._etot[:, :self.batch].mean(0).cpu().numpy()
            for i, b in enumerate(self.baths):
                d["curmean%d" % i] = b._cur_dev[:, :self.batch].mean(0).cpu().numpy()
        else:
            d["energy"] = self.etot
            if self._savepq:
                d["ps"], d["qs"] = self.ps, self.qs
            for i, b in enumerate(self.baths):
                d["noise%d" % i] = b.noise
                d["cur%d" % i] = b.cur
                if self.record_fhis and i in self._fhis:
                    d["fhis%d" % i] = self._host(self._fhis[i])         # (nmd, nc[, batch]): the bath's dofs only
        name = self._ckpt_name(id)
        tmp = name + ".tmp" + os.path.splitext(name)[1]
        if self.batch == 1:
            self._write_nc(tmp, d)
        else:
            d["created"] = N.array([time.time()])
            N.savez(tmp, **d)
        os.replace(tmp, name)

    def _write_nc(self, path, d):
        """The reference's MD<j>.nc layout: dimensions nph, one, two, mem, nmd, nnmd, n<i>; variables
        noise<i> (nnmd, n<i>), [fhis<i> (nnmd, nph) when recorded], ps, qs, power, power2, energy, p, q, t,
        ipie, phis, qhis (only row 0 of qhis is ever read, md.py:380), plus cur<i> (nnmd) for the restart."""
        from scipy.io import netcdf_file
        f = netcdf_file(path, "w")
        try:
            f.title = "Output from md.py"
            f.history = "Created " + time.ctime(time.time())
            for name, n in (("nph", self.nph), ("one", 1), ("two", 2), ("mem", max(self.ml, 1)), ("nmd", self.nmd),
                            ("nnmd", self.nmd)):
                f.createDimension(name, n)
            for i, b in enumerate(self.baths):
                f.createDimension("n%d" % i, b.nc)

            def put(label, a, dims):
                v = f.createVariable(label, "d", dims)
                v[:] = N.asarray(a, dtype=float).reshape([f.dimensions[k] for k in dims])
                v.units = ""
            for i, b in enumerate(self.baths):
                put("noise%d" % i, d["noise%d" % i], ("nnmd", "n%d" % i))
                put("cur%d" % i, d["cur%d" % i], ("nnmd",))
                if "fhis%d" % i in d:
                    full = N.zeros((self.nmd, self.nph))                # the reference's (nnmd, nph) layout
                    full[:, N.asarray(b.cids)] = N.asarray(d["fhis%d" % i]).reshape(self.nmd, b.nc)
                    put("fhis%d" % i, full, ("nnmd", "nph"))
            if "ps" in d:
                put("ps", d["ps"], ("nnmd", "nph"))
                put("qs", d["qs"], ("nnmd", "nph"))
            put("power", d["power"], ("nnmd", "two"))
            put("power2", d["power2"], ("nnmd", "two"))
            put("energy", d["energy"], ("nnmd",))
            put("p", d["p"], ("nph",))
            put("q", d["q"], ("nph",))
            put("t", [float(self.t)], ("one",))
            put("ipie", d["ipie"], ("one",))
            phis = N.asarray(d["phis"], dtype=float).reshape(-1, self.nph)
            mem = N.zeros((max(self.ml, 1), self.nph))
            mem[:min(len(phis), len(mem))] = phis[:len(mem)]
            put("phis", mem, ("mem", "nph"))
            qh = N.zeros_like(mem)
            qh[0] = N.asarray(d["q"], dtype=float).ravel()
            put("qhis", qh, ("mem", "nph"))
        finally:
            f.close()

    @staticmethod
    def _read_ckpt(path):
        """npz or NetCDF-3 checkpoint -> dict of arrays (ipie and t as 1-element arrays)."""
        if path.endswith(".npz"):
            return N.load(path)
        from scipy.io import netcdf_file
        f = netcdf_file(path, "r", mmap=False)
        try:
            return {k: N.array(v[...], dtype=float) for k, v in f.variables.items()}
        finally:
            f.close()

    def _resume(self, j):
        """md.py:541-582: resume an unfinished segment, skip a finished one, or chain the
        state from the previous segment's checkpoint."""
        fn, fnm = self._ckpt_name(j), self._ckpt_name(j - 1)
        if self.write_files and os.path.isfile(fn):
            d = self._read_ckpt(fn)
            ipie = int(d["ipie"][0])
            if ipie + 1 < self.npie:
                self.p, self.q = d["p"], d["q"]
                self.t = int(d["t"][0])
                self.set_history(d["phis"])
                self.power, self.power2 = d["power"], d["power2"]
                for i, b in enumerate(self.baths):
                    b.noise = d["noise%d" % i]
                    b._ngnoi = j + 1
                    self._upload(b._alloc_cur(), d["cur%d" % i])
                    if self.record_fhis and i in self._fhis and "fhis%d" % i in d:
                        fh = N.asarray(d["fhis%d" % i])
                        if fh.shape[:2] == (self.nmd, self.nph):        # MD<j>.nc stores the scattered layout
                            fh = fh[:, N.asarray(b.cids)]
                        self._upload(self._fhis[i], fh)
                self.ResetSavepq(self._savepq)
                if self._savepq and "ps" in d:
                    self._upload(self._ps, d["ps"])
                    self._upload(self._qs, d["qs"])
                return ipie
            if ipie + 1 == self.npie:
                self.power, self.power2 = d["power"], d["power2"]
                self.t = int(d["t"][0])
                return "finished"
            raise ValueError("ipie error")
        if self.write_files and os.path.isfile(fnm):
            d = self._read_ckpt(fnm)
            self.p, self.q = d["p"], d["q"]
            self.t = int(d["t"][0])
            if d["phis"].shape[:2] == (self.ml, self.nph):
                self.set_history(d["phis"])
        return None


def sameq(q1, q2):
    """md.py:724-736."""
    if len(q1) != len(q2):
        return False
    return bool(N.max(N.abs(N.asarray(q1) - N.asarray(q2))) < 10e-10)


def ApplyConstraint(f, constr=None):
    """Remove the components of a host vector f along each normalised constraint vector
    (md.py:739-762); all dot products use the un-projected f."""
    if constr is None:
        return f
    f = N.array(f, dtype=float)
    c = N.array(constr, dtype=float)
    nf = 1.0 * f
    for i in range(c.shape[0]):
        nn = LA.norm(c[i])
        ci = c[i] / nn if nn != 0 else c[i]
        nf = nf - N.dot(f, ci) * ci
    return nf
