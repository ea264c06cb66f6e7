"""Scalar helpers of the reference (functions.py:26-101), vectorised for table building.

``flinterp`` keeps the reference's interpolation convention (functions.py:61-78): the slope
term is taken with the sign that mirrors the value about the nearest grid point; it is exact
on grid points and constant below the first / above the last cell midpoint.  Every table
that feeds the device path (friction kernel, noise covariances) is built through it, so the
device results follow the reference's tables, not a "corrected" interpolation.
"""
import numpy as N

from . import units as U


def bose(w, T):
    """Bose-Einstein occupation with the reference's special cases (functions.py:26-45):
    T == 0: 0 for w > 0, -1 for w < 0, 1/(exp(1/kb)-1) at w == 0;  T > 0: 0 at w == 0."""
    w = N.asarray(w, dtype=float)
    out = N.zeros_like(w)
    if T == 0.0:
        out = N.where(w < 0.0, -1.0, 0.0)
        with N.errstate(over="ignore"):
            out = N.where(w == 0.0, 1.0 / (N.exp(1.0 / U.kb) - 1.0), out)
        return out if out.ndim else float(out)
    nz = w != 0.0
    with N.errstate(over="ignore", divide="ignore"):
        val = 1.0 / (N.exp(N.where(nz, w, 1.0) / U.kb / T) - 1.0)
    out = N.where(nz, val, 0.0)
    return out if out.ndim else float(out)


def fermi(ep, mu, T):
    """Fermi function (functions.py:47-59)."""
    if T == 0.0:
        return 1.0 if ep < mu else (0.0 if ep > mu else 0.5)
    return 1 / (N.exp((ep - mu) / U.kb / T) + 1)


def nearest(b, bs):
    """Index of the first element of bs closest to b (functions.py:80-86)."""
    return int(N.argmin(N.abs(N.asarray(bs) - b)))


def flinterp(x, xs, ys):
    """Reference-convention interpolation of the table (xs, ys) at x; x may be an array,
    in which case the result has shape x.shape + ys.shape[1:]."""
    xs = N.asarray(xs, dtype=float)
    ys = N.asarray(ys)
    xa = N.atleast_1d(N.asarray(x, dtype=float))
    n = len(xs)
    idx = N.argmin(N.abs(xs[None, :] - xa[:, None]), axis=1)
    out = ys[idx].copy()
    if n > 1:
        inner = (idx > 0) & (idx < n - 1)
        if inner.any():
            ii = idx[inner]
            dd = xs[ii] - xa[inner]
            other = N.where(dd > 0, ii - 1, ii + 1)
            coef = dd / (xs[ii] - xs[other])
            shape = (-1,) + (1,) * (ys.ndim - 1)
            out[inner] = ys[ii] + coef.reshape(shape) * (ys[ii] - ys[other])
    return out[0] if N.ndim(x) == 0 else out


def flinterp_weights(x, xs):
    """flinterp as two table rows and weights: flinterp(x, xs, ys) = w0 * ys[i0] + w1 * ys[i1] (up to
    rounding), with i0 the nearest grid point, i1 its neighbour on the side flinterp uses and
    (w0, w1) = (1 + c, -c) in the reference's mirrored-slope convention (functions.py:61-78).
    Returns int32 (n, 2) indices and float (n, 2) weights; lets a kernel assemble interpolated tables
    without materialising them (pymd_eigh_factors)."""
    xs = N.asarray(xs, dtype=float)
    xa = N.atleast_1d(N.asarray(x, dtype=float))
    n = len(xs)
    i0 = N.argmin(N.abs(xs[None, :] - xa[:, None]), axis=1)
    i1 = i0.copy()
    c = N.zeros(len(xa))
    if n > 1:
        inner = (i0 > 0) & (i0 < n - 1)
        ii = i0[inner]
        dd = xs[ii] - xa[inner]
        other = N.where(dd > 0, ii - 1, ii + 1)
        i1[inner] = other
        c[inner] = dd / (xs[ii] - xs[other])
    return N.stack([i0, i1], axis=1).astype(N.int32), N.stack([1.0 + c, -c], axis=1)


def rpadleft(bs, b):
    """Shift ``b`` in at the front, drop the last row (functions.py:88-95).  Host utility;
    on the device the history is a ring buffer (pymd_bind_history)."""
    bs = N.asarray(bs)
    if len(bs) < 1:
        raise ValueError("len(bs) is less than 1")
    return N.concatenate((N.asarray([b]), bs[:-1]), axis=0)


def mm(*args):
    out = N.asarray(args[0]).copy()
    for a in args[1:]:
        out = N.dot(out, a)
    return out
