"""Displacement / velocity power spectra (reference spectrum.py:8-45) on the GPU."""
import numpy as N

from . import _lib

# transform at most this many (dof x trajectory) columns at a time (bounds the FFT scratch)
MAX_CHUNK_BYTES = 8 << 30


def ensemble_powerspec(traj_dev, dt, nmd, nvalid, wsq, host=True):
    """sum over dofs and over the first ``nvalid`` trajectories of |Fourier1D(x)|^2/(dt nmd)
    (times w_i^2 when wsq) for a device array [nmd][nd][ldb]; returns a host vector (nmd,)
    (``host=False``: the device tensor, no synchronisation)."""
    import torch
    lib = _lib.load()
    n, nd, ldb = traj_dev.shape
    if n != nmd:
        raise ValueError("power: shape error!")
    total = nd * ldb
    chunk = min(total, max(32, (MAX_CHUNK_BYTES // (16 * nmd)) // 32 * 32))
    out = torch.empty(nmd, dtype=torch.float64, device=traj_dev.device)
    work = torch.empty(lib.pymd_powerspec_work_bytes(nmd, nd, ldb, chunk), dtype=torch.uint8,
                       device=traj_dev.device)
    if traj_dev.stride(2) != 1 or traj_dev.stride(1) != ldb:
        raise ValueError("power: the rows of the saved trajectories must be contiguous [nd][ldb] blocks")
    if traj_dev.stride(0) == nd * ldb:
        _lib.call("pymd_powerspec", traj_dev.data_ptr(), nmd, nd, ldb, int(nvalid), float(dt), 1 if wsq else 0,
                  out.data_ptr(), work.data_ptr(), chunk, _lib.stream_ptr())
    else:                                               # rows of an overlaid record buffer (md.enable_record_overlay)
        _lib.call("pymd_powerspec_ld", traj_dev.data_ptr(), int(traj_dev.stride(0)), nmd, nd, ldb, int(nvalid), float(dt),
                  1 if wsq else 0, out.data_ptr(), work.data_ptr(), chunk, _lib.stream_ptr())
    return out.cpu().numpy() if host else out


def time_means(rec_dev, out=None):
    """Per-trajectory time mean of a device record [rows][ldb] (bath.cur, md.etot): device vector [ldb]."""
    import torch
    lib = _lib.load()
    rows, ldb = rec_dev.shape
    if out is None:
        out = torch.empty(ldb, dtype=torch.float64, device=rec_dev.device)
    work = torch.empty(lib.pymd_column_sums_work_bytes(ldb), dtype=torch.uint8, device=rec_dev.device)
    _lib.call("pymd_column_sums", rec_dev.data_ptr(), rows, ldb, 1.0 / rows, out.data_ptr(), work.data_ptr(),
              _lib.stream_ptr())
    return out


def _single(xs, dt, nmd, wsq):
    import torch
    a = N.asarray(xs, dtype=N.float64)
    if a.ndim != 2 or a.shape[0] != nmd:
        raise ValueError("power: shape error!")
    dev = torch.empty((nmd, a.shape[1], 32), dtype=torch.float64, device="cuda").zero_()
    dev[:, :, 0] = torch.as_tensor(a).cuda()
    s = ensemble_powerspec(dev, dt, nmd, 1, wsq)
    dw = 2. * N.pi / dt / nmd
    return N.stack((dw * N.arange(nmd), s), axis=1)


def powerspec(qs, dt, nmd):
    """rows (w_i, w_i^2 sum_dof |q(w_i)|^2/(dt nmd)) for one trajectory qs (nmd, nph)
    (spectrum.py:8-25)."""
    return _single(qs, dt, nmd, True)


def powerspec2(ps, dt, nmd):
    """rows (w_i, sum_dof |p(w_i)|^2/(dt nmd))  (spectrum.py:28-45)."""
    return _single(ps, dt, nmd, False)
