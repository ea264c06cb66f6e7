"""Coloured quantum noise (reference noise.py).

Host side (setup, identical for every trajectory of the ensemble, done with numpy exactly
as the reference does): the noise covariance per positive frequency and its
eigen-decomposition, packed as spectral factors L_f = U_f diag(sqrt(max(lambda_f, 0))).

Device side (per trajectory, libpymd_b200): white noise (Philox or supplied), X_f = L_f xi_f,
Hermitian mirror, w->t FFT, real part  (noise.py:61-88, 156-190, vargau 252-284).
"""
import numpy as N

from . import units as U
from . import _lib
from .functions import bose, flinterp, flinterp_weights
from .matrix import chkShape, hermitianize


def equ(w, cut, T, classical=False, zpmotion=True):
    """2 hw (n_B(hw,T) + zp) for hw < cut, else 0; 2 kb T in the classical limit and at
    hw == 0 (noise.py:229-250).  ``w`` may be an array."""
    hw = U.hbar * N.asarray(w, dtype=float)
    zp = 0.5 if zpmotion is True else 0.0
    if classical:
        val = N.full_like(hw, 2.0 * U.kb * T)
    else:
        val = N.where(hw == 0, 2.0 * U.kb * T, 2.0 * hw * (zp + bose(hw, T)))
    out = N.where(hw < cut, val, 0.0)
    return out if out.ndim else float(out)


def phnoisew(gamma, wl, T, phcut, classical=False, zpmotion=True):
    """Phonon noise spectrum equ(w) gamma(w) on the wl grid (noise.py:16-34)."""
    gamma = N.asarray(gamma)
    e = equ(N.asarray(wl), phcut, T, classical, zpmotion)
    return e.reshape((-1,) + (1,) * (gamma.ndim - 1)) * gamma


def _electron_cov(w, efric, exim, exip, bias, T, ecut, classical, zpmotion, delta=1.0):
    """noise.py:103-127 / 156-170: equilibrium part + the two bias-shifted parts."""
    w = N.asarray(w, dtype=float)
    efric, exim, exip = N.asarray(efric), N.asarray(exim), N.asarray(exip)
    aw = (delta * equ(w, ecut, T, classical, zpmotion))[:, None, None]
    awm = (delta * equ(U.hbar * w - bias, ecut, T, classical, zpmotion))[:, None, None]
    awp = (delta * equ(U.hbar * w + bias, ecut, T, classical, zpmotion))[:, None, None]
    amat = aw * efric + (-0.5 * aw * exip + 0.5 * awm * (exip + 1j * exim)) \
        + (-0.5 * aw * exip + 0.5 * awp * (exip - 1j * exim))
    return 0.5 * (amat + N.conj(N.transpose(amat, (0, 2, 1))))


def enoisew(wl, efric, exim, exip, bias, T, ecut, classical=False, zpmotion=True):
    """Electron noise spectrum on the wl grid, no sampling (noise.py:91-129)."""
    nc = chkShape(efric)
    if chkShape(exim) != nc or chkShape(exip) != nc:
        raise ValueError("enoisew: efric shape error!")
    return _electron_cov(wl, efric, exim, exip, bias, T, ecut, classical, zpmotion)


def _batched_eigh(amat):
    """numpy.linalg.eigh of a stack of Hermitian matrices, the stack split over host threads (LAPACK
    releases the GIL; every matrix is decomposed by the same routine as in one big call, so the
    result is bit-identical to N.linalg.eigh(amat), noise.py:69,172)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    nf = amat.shape[0]
    nt = min(os.cpu_count() or 1, 16, max(1, nf // 256))
    if nt <= 1:
        return N.linalg.eigh(amat)
    bounds = N.linspace(0, nf, nt + 1).astype(int)
    with ThreadPoolExecutor(nt) as ex:
        parts = list(ex.map(lambda i: N.linalg.eigh(amat[bounds[i]:bounds[i + 1]]), range(nt)))
    return N.concatenate([p[0] for p in parts]), N.concatenate([p[1] for p in parts])


class SpectralFactors:
    """lam[f, i], U[f, :, i] (as numpy.linalg.eigh returns them) and L = U sqrt(max(lam,0))."""

    def __init__(self, lam, vec):
        self.lam, self.U = lam, vec
        self.L = vec * N.sqrt(N.maximum(lam, 0.0))[:, None, :]
        self.complex = N.iscomplexobj(vec)
        self.shape = self.L.shape
        self._dev = {}                      # device -> (lre, lim): uploaded once, reused by every gnoi()
        self._host = None

    def _pinned(self):
        """L as page-locked host tensors (real, imaginary), made once: uploads then run at PCIe/NVLink-C2C
        speed and can be asynchronous."""
        import torch
        if self._host is None:
            re = torch.as_tensor(N.ascontiguousarray(self.L.real), dtype=torch.float64)
            im = torch.as_tensor(N.ascontiguousarray(self.L.imag), dtype=torch.float64) if self.complex else None
            try:
                re = re.pin_memory()
                im = im.pin_memory() if im is not None else None
            except RuntimeError:
                pass
            self._host = (re, im)
        return self._host

    def device_factors(self, dev):
        """L on the device (real and imaginary parts as separate [nf][nc][nc] arrays)."""
        key = str(dev)
        if key not in self._dev:
            re, im = self._pinned()
            lre = re.to(dev, non_blocking=True)
            lim = im.to(dev, non_blocking=True) if im is not None else None
            self._dev[key] = (lre, lim)
        return self._dev[key]

    def drop_device_copy(self):
        """Forget the device copies: the next gnoi() uploads the factors again."""
        self._dev = {}

    def upload_copy(self, dev):
        """A fresh device copy of L (asynchronous, from pinned memory, on the current stream) that is NOT yet the
        one gnoi() uses: ensemble.run_pipelined uploads the next sub-ensemble's inputs on a copy stream."""
        re, im = self._pinned()
        return (re.to(dev, non_blocking=True), im.to(dev, non_blocking=True) if im is not None else None)

    def adopt_device_copy(self, dev, pair):
        self._dev[str(dev)] = pair

    def as_oracle(self):
        return self.lam, self.U


class DeviceSpectralFactors:
    """Spectral factors decomposed on the GPU (pymd_eigh_factors, batched Jacobi): the covariances are
    uploaded once, L stays on the device where pymd_noise_generate reads it, and lam / U / L are copied
    to the host only when asked for.  Eigenvectors carry an arbitrary phase, so L equals the LAPACK
    factor up to a unitary diagonal: the noise has the same covariance L L^H, not the same realisation
    for a given xi -- use the host factors (default) to reproduce reference noise element by element."""

    def __init__(self, amat, device, keep_vectors=False, combination=None):
        """``amat``: the (nf, nc, nc) covariances, or -- with ``combination=(idx, coef)`` -- a small table of
        matrices from which the kernel assembles C_f = herm(sum_m coef[f, m] amat[idx[f, m]])."""
        import torch
        amat = N.asarray(amat)
        nc = amat.shape[1]
        self.complex = bool(N.iscomplexobj(amat))
        self.device = torch.device(device)
        kw = dict(dtype=torch.float64, device=self.device)
        cre = torch.as_tensor(N.ascontiguousarray(amat.real), **kw)
        cim = torch.as_tensor(N.ascontiguousarray(amat.imag), **kw) if self.complex else None
        idx = coef = None
        nf, ncomb = amat.shape[0], 0
        if combination is not None:
            idx = torch.as_tensor(N.ascontiguousarray(combination[0], dtype=N.int32), device=self.device)
            coef = torch.as_tensor(N.ascontiguousarray(combination[1], dtype=float), **kw)
            nf, ncomb = idx.shape
            if tuple(coef.shape) != (nf, ncomb) or int(idx.max()) >= amat.shape[0] or int(idx.min()) < 0:
                raise ValueError("DeviceSpectralFactors: bad combination")
        self.shape = (nf, nc, nc)
        self._lam = torch.empty((nf, nc), **kw)
        self._lre = torch.empty((nf, nc, nc), **kw)
        self._lim = torch.empty((nf, nc, nc), **kw) if self.complex else None
        self._ure = torch.empty((nf, nc, nc), **kw) if keep_vectors else None
        self._uim = torch.empty((nf, nc, nc), **kw) if keep_vectors and self.complex else None
        self._sweeps = torch.empty(nf, dtype=torch.int32, device=self.device)
        wb = _lib.load().pymd_eigh_work_bytes(nf, nc, 1 if self.complex else 0)
        work = torch.empty(wb, dtype=torch.uint8, device=self.device) if wb else None
        with torch.cuda.device(self.device):
            _lib.call("pymd_eigh_factors", cre.data_ptr(), _lib.ptr(cim), _lib.ptr(idx), _lib.ptr(coef), ncomb, nf, nc,
                      self._lam.data_ptr(), self._lre.data_ptr(), _lib.ptr(self._lim), _lib.ptr(self._ure),
                      _lib.ptr(self._uim), self._sweeps.data_ptr(), _lib.ptr(work), _lib.stream_ptr())
            if work is not None:
                torch.cuda.current_stream().synchronize()      # the scratch is freed on return

    @staticmethod
    def supported(amat):
        return amat.shape[1] <= _lib.load().pymd_eigh_max_n(1 if N.iscomplexobj(amat) else 0)

    def _host(self, re, im):
        a = re.cpu().numpy()
        return a + 1j * im.cpu().numpy() if im is not None else a

    @property
    def lam(self):
        return self._lam.cpu().numpy()

    @property
    def L(self):
        return self._host(self._lre, self._lim)

    @property
    def U(self):
        if self._ure is None:
            raise ValueError("eigenvectors were not kept (keep_vectors=False)")
        return self._host(self._ure, self._uim)

    @property
    def sweeps(self):
        return self._sweeps.cpu().numpy()

    def device_factors(self, dev):
        import torch
        if torch.device(dev) == self.device:
            return self._lre, self._lim
        return self._lre.to(dev), (self._lim.to(dev) if self._lim is not None else None)

    def drop_device_copy(self):
        """The factors live on the device; there is no host copy to upload again."""

    def upload_copy(self, dev):
        return self.device_factors(dev)

    def adopt_device_copy(self, dev, pair):
        pass

    def as_oracle(self):
        return self.lam, self.U


def setup_on_device():
    """PYMD_SETUP_DEVICE=1: decompose the noise covariances (and build gmem tables) on the GPU."""
    import os
    return os.environ.get("PYMD_SETUP_DEVICE", "0") == "1"


def _device_ok(tables, device, strict):
    """``strict``: raise when the matrices do not fit the device kernel (an explicit request); otherwise
    (PYMD_SETUP_DEVICE=1) such baths keep the LAPACK path."""
    if device is None:
        return False
    if DeviceSpectralFactors.supported(tables):
        return True
    if strict:
        raise ValueError("noise: nc=%d is too large for the on-device eigen-decomposition" % tables.shape[1])
    return False


def _decompose(amat, device, strict=True, keep_vectors=False):
    if _device_ok(amat, device, strict):
        return DeviceSpectralFactors(amat, device, keep_vectors)
    lam, vec = _batched_eigh(amat)
    return SpectralFactors(lam, vec)


def phonon_factors(gamma, wl, T, phcut, dt, nmd, classical=False, zpmotion=True, device=None, strict=True,
                   keep_vectors=False):
    """Covariance delta*equ(w_f)*gamma~(w_f), f = 0..nmd/2 (noise.py:61-66), eigh (noise.py:69);
    ``device``: decompose on that GPU instead of LAPACK."""
    nf = int(nmd / 2) + 1
    dw = 2.0 * N.pi / dt / nmd
    delta = dt * nmd
    w = dw * N.arange(nf)
    gamma = N.asarray(gamma)
    if _device_ok(gamma, device, strict):
        # the kernel assembles delta equ(w_f) [w0 gamma[i0] + w1 gamma[i1]] itself: nothing of size nf x nc x nc
        # is built or uploaded
        idx, wt = flinterp_weights(w, wl)
        coef = (delta * equ(w, phcut, T, classical, zpmotion))[:, None] * wt
        return DeviceSpectralFactors(gamma, device, keep_vectors, combination=(idx, coef))
    amat = (delta * equ(w, phcut, T, classical, zpmotion))[:, None, None] * flinterp(w, wl, gamma)
    amat = 0.5 * (amat + N.conj(N.transpose(amat, (0, 2, 1))))
    return _decompose(amat, device, strict, keep_vectors)


def electron_factors(efric, exim, exip, bias, T, ecut, dt, nmd, classical=False, zpmotion=True, device=None,
                     strict=True, keep_vectors=False):
    """Bias-dependent Hermitian covariance (noise.py:156-170), eigh (noise.py:172)."""
    nf = int(nmd / 2) + 1
    dw = 2.0 * N.pi / dt / nmd
    w = dw * N.arange(nf)
    tables = N.array([N.asarray(efric) + 0j, N.asarray(exip) + 0j, 1j * N.asarray(exim)])
    if _device_ok(tables, device, strict):
        # _electron_cov regrouped: aw efric + [(awm + awp)/2 - aw] exip + (awm - awp)/2 (i exim)
        delta = dt * nmd
        aw = delta * equ(w, ecut, T, classical, zpmotion)
        awm = delta * equ(U.hbar * w - bias, ecut, T, classical, zpmotion)
        awp = delta * equ(U.hbar * w + bias, ecut, T, classical, zpmotion)
        coef = N.stack([aw, 0.5 * (awm + awp) - aw, 0.5 * (awm - awp)], axis=1)
        idx = N.tile(N.arange(3, dtype=N.int32), (nf, 1))
        return DeviceSpectralFactors(tables, device, keep_vectors, combination=(idx, coef))
    amat = _electron_cov(w, efric, exim, exip, bias, T, ecut, classical, zpmotion, delta=dt * nmd)
    return _decompose(amat, device, strict, keep_vectors)


# ------------------------------------------------------------------ device pipeline
MAX_WORK_BYTES = 16 << 30    # scratch bound (FFT ping-pong + white-noise chunk) for noise generation

_SCRATCH = {}


def _work_budget(device):
    """Scratch bytes one generate call may use: MAX_WORK_BYTES, but no more than a third of the memory that is free
    (counting the scratch already held), at least 1 GiB.  Larger trajectory chunks mean fewer re-reads of the spectral
    factors and longer column loops per CTA (cfg2, 4096 trajectories: 193.6 / 176.6 / 171.1 ms at 4 / 8 / 16 GiB)."""
    import torch
    free, _ = torch.cuda.mem_get_info(device)
    held = _SCRATCH.get(str(device))
    free += held.numel() if held is not None else 0
    return int(min(MAX_WORK_BYTES, max(1 << 30, free // 3)))


def _scratch(nbytes, device):
    """Grow-only FFT scratch per device, shared by all baths (the generate calls are stream-ordered)."""
    import torch
    key = str(device)
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < nbytes:
        _SCRATCH[key] = None
        buf = None
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _SCRATCH[key] = buf
    return buf


def white_noise(nf, nc, ldb, seed, stream_id, traj0=0, device=None):
    """Standard normals [nf][nc][ldb] from the library's Philox generator (replaces the
    sequential N.random.normal draws of noise.vargau, noise.py:282)."""
    import torch
    xi = torch.empty((nf, nc, ldb), dtype=torch.float64, device=device or "cuda")
    _lib.call("pymd_philox_normal", xi.data_ptr(), nf * nc, ldb, int(seed) & (2 ** 64 - 1),
              int(stream_id) & 0xFFFFFFFF, int(traj0), _lib.stream_ptr())
    return xi


def generate(factors, xi, dt, nmd, out=None):
    """noise[t][i][traj] on the device from spectral factors and white noise xi [nf][nc][ldb]."""
    import torch
    nf, nc, ldb = xi.shape
    if nf != int(nmd / 2) + 1 or tuple(factors.shape) != (nf, nc, nc):
        raise ValueError("noise.generate: shape mismatch")
    dev = xi.device
    lre, lim = factors.device_factors(dev)
    if out is None:
        out = torch.empty((nmd, nc, ldb), dtype=torch.float64, device=dev)
    if not out.is_contiguous():
        raise ValueError("noise.generate: injected white noise needs a contiguous noise array (no record overlay)")
    # trajectories are transformed in chunks so that the FFT scratch stays below MAX_WORK_BYTES
    per_col = _lib.load().pymd_noise_work_bytes(nmd, nc, 1)
    cw = max(32, min(ldb, (_work_budget(dev) // per_col) // 32 * 32))
    work = _scratch(per_col * cw, dev)
    for n0 in range(0, ldb, cw):
        n1 = min(ldb, n0 + cw)
        _lib.call("pymd_noise_generate", lre.data_ptr(), _lib.ptr(lim), xi.data_ptr(), nmd, nc, ldb, n0, n1,
                  float(dt), out.data_ptr(), work.data_ptr(), _lib.stream_ptr())
    return out


def generate_philox(factors, dt, nmd, ldb, batch, seed, stream_id, traj0, out):
    """noise[t][i][traj] from spectral factors and the library's Philox normals, one trajectory chunk at a time:
    the white noise (nf x nc x ldb doubles, 16 GB for the 60-dof bath of 2048 trajectories) never exists as a
    whole.  Element (f, i, traj0 + n) of the white noise is the same whatever the chunking or the sharding."""
    import torch
    nf, nc, _ = factors.shape
    dev = out.device
    if out.stride(2) != 1 or out.stride(1) != ldb:
        raise ValueError("noise.generate_philox: the output rows must be contiguous [nc][ldb] blocks")
    lre, lim = factors.device_factors(dev)
    per_col = _lib.load().pymd_noise_work_bytes(nmd, nc, 1) + 8 * nf * nc
    cw = max(32, min(ldb, (_work_budget(dev) // per_col) // 32 * 32))
    work = _scratch(per_col * cw, dev)
    wbytes = _lib.load().pymd_noise_work_bytes(nmd, nc, cw)
    xi = work[wbytes:wbytes + 8 * nf * nc * cw].view(torch.float64)
    for n0 in range(0, ldb, cw):
        w = min(cw, ldb - n0)
        x = xi[:nf * nc * w].view(nf, nc, w)
        _lib.call("pymd_philox_normal", x.data_ptr(), nf * nc, w, int(seed) & (2 ** 64 - 1),
                  int(stream_id) & 0xFFFFFFFF, int(traj0) + n0, _lib.stream_ptr())
        if n0 + w > batch:
            x[:, :, max(batch - n0, 0):] = 0.0          # padding trajectories stay at rest
        if out.stride(0) == nc * ldb:
            _lib.call("pymd_noise_generate_chunk", lre.data_ptr(), _lib.ptr(lim), x.data_ptr(), w, nmd, nc, ldb, n0,
                      float(dt), out.data_ptr(), work.data_ptr(), _lib.stream_ptr())
        else:                                           # rows of an overlaid record buffer (md.enable_record_overlay)
            _lib.call("pymd_noise_generate_chunk_ld", lre.data_ptr(), _lib.ptr(lim), x.data_ptr(), w, nmd, nc, ldb, n0,
                      float(dt), out.data_ptr(), int(out.stride(0)), work.data_ptr(), _lib.stream_ptr())
    return out


def _as_device_xi(xi, nf, nc, batch, ldb, device):
    """Accept user-supplied normals shaped (nf, nc) or (nf, nc, batch); pad to ldb."""
    import torch
    x = torch.as_tensor(N.asarray(xi) if not hasattr(xi, "device") else xi, dtype=torch.float64)
    if x.dim() == 2:
        x = x[:, :, None]
    if x.shape[0] != nf or x.shape[1] != nc or x.shape[2] != batch:
        raise ValueError("xi must have shape (%d, %d[, %d])" % (nf, nc, batch))
    full = torch.zeros((nf, nc, ldb), dtype=torch.float64, device=device)
    full[:, :, :batch] = x.to(device)
    return full


def phnoise(gamma, wl, T, phcut, dt, nmd, classical=False, zpmotion=True, xi=None, seed=0):
    """Single-trajectory convenience with the reference's signature (noise.py:38): returns a
    host array (nmd, nc).  Runs on the GPU."""
    import torch
    fac = phonon_factors(gamma, wl, T, phcut, dt, nmd, classical, zpmotion)
    nf, nc = fac.lam.shape
    x = white_noise(nf, nc, 32, seed, 0) if xi is None else _as_device_xi(xi, nf, nc, 1, 32, "cuda")
    return generate(fac, x, dt, nmd)[:, :, 0].cpu().numpy()


def enoise(efric, exim, exip, bias, T, ecut, dt, nmd, classical=False, zpmotion=True, xi=None, seed=0):
    """Single-trajectory convenience with the reference's signature (noise.py:135)."""
    fac = electron_factors(efric, exim, exip, bias, T, ecut, dt, nmd, classical, zpmotion)
    nf, nc = fac.lam.shape
    x = white_noise(nf, nc, 32, seed, 0) if xi is None else _as_device_xi(xi, nf, nc, 1, 32, "cuda")
    return generate(fac, x, dt, nmd)[:, :, 0].cpu().numpy()
