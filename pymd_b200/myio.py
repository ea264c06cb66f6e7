"""NetCDF inputs of the reference scripts (reference myio.py:146-360): the electron-phonon
file ``EPH.nc`` (Wlist, hw, U, DynMat, Re/ImSigL, Re/ImSigR, Friction, NC, NCP, zeta1, zeta2),
PHrun's ``Dev*.nc`` (hw, U), ``Sig.nc``, ``MD.nc`` geometry and the bias-dependent coupling
file.  Same function names, argument order and returned fields as the reference, so that
``from myio import *`` in a driver script can become ``from pymd_b200.myio import *``.

The reference opens the files with netCDF4 / Scientific.IO.NetCDF, neither of which exists in
this image; ``scipy.io.netcdf_file`` is used instead.  It reads and writes NetCDF-3 (classic and
64-bit offset) files, the format Scientific.IO.NetCDF produces; HDF5-based NetCDF-4 files must be
converted first (``nccopy -k classic``).  Host-side set-up only, not on the stepping path.
Errors are exceptions (the reference prints and calls sys.exit()).
"""
import numpy as N
from scipy.io import netcdf_file

__all__ = ["ReadEPHNCFile", "ReadNewEPHNCFile", "WriteEPHNCfile", "Write2NetCDFFile", "ReadNetCDFVar",
           "ReadMDNCFile", "ReadDynmat", "ord2idx", "ReadSig", "ReadwbLambda", "ReadLambda"]


class _Record(object):
    """Attribute bag, as the reference's throw-away ``class eph: pass``."""

    def __init__(self, filename=None):
        if filename is not None:
            self.filename = filename


def _open(filename):
    try:
        return netcdf_file(filename, "r", mmap=False)
    except TypeError as e:          # scipy's message for an HDF5 container
        raise IOError("%s is not a NetCDF-3 file (%s); convert NetCDF-4/HDF5 files with "
                      "'nccopy -k classic'" % (filename, e))


def _var(f, name):
    if name not in f.variables:
        raise KeyError("%s: no variable '%s' (has: %s)" % (getattr(f, "filename", "file"), name,
                                                          ", ".join(sorted(f.variables))))
    return N.array(f.variables[name][...], dtype=float)


def _cplx(f, stem):
    return _var(f, "Re" + stem) + 1j * _var(f, "Im" + stem)


def ReadEPHNCFile(filename):
    """Dynamical matrix, lead self-energies and electronic friction/NC matrices (myio.py:146-170)."""
    f = _open(filename)
    try:
        eph = _Record(filename)
        eph.wl, eph.hw, eph.U, eph.DynMat = (_var(f, k) for k in ("Wlist", "hw", "U", "DynMat"))
        eph.SigL, eph.SigR = _cplx(f, "SigL"), _cplx(f, "SigR")
        eph.efric, eph.xim, eph.xip = _var(f, "Friction"), _var(f, "NC"), _var(f, "NCP")
    finally:
        f.close()
    return eph


def ReadNewEPHNCFile(filename):
    """ReadEPHNCFile plus the bias-linear terms zeta1, zeta2 (myio.py:173-197)."""
    eph = ReadEPHNCFile(filename)
    f = _open(filename)
    try:
        eph.zeta1, eph.zeta2 = _var(f, "zeta1"), _var(f, "zeta2")
    finally:
        f.close()
    return eph


def Write2NetCDFFile(file, var, varLabel, dimensions, units=None, description=None):
    """One double-precision variable into an open file (myio.py:231-236)."""
    v = file.createVariable(varLabel, "d", dimensions)
    v[:] = N.asarray(var, dtype=float)
    if units:
        v.units = units
    if description:
        v.description = description


def WriteEPHNCfile(filename, wl, hw, U, DynMat, SigL, SigR, Friction, NC, NCP, zeta1, zeta2):
    """Write the EPH.nc schema (myio.py:199-228).  The reference declares ReSigR/ImSigR with the
    LEFT lead's dimension 'Nsl' (myio.py:220-221), which fails for unequal leads; here the right
    lead gets its own dimension 'Nsr' (the reference creates it, but never uses it)."""
    SigL, SigR = N.asarray(SigL), N.asarray(SigR)
    fn = netcdf_file(filename, "w")
    try:
        fn.createDimension("NPh", len(hw))
        fn.createDimension("NWl", len(wl))
        fn.createDimension("Nsl", SigL.shape[1])
        fn.createDimension("Nsr", SigR.shape[1])
        Write2NetCDFFile(fn, wl, "Wlist", ("NWl",), units="eV")
        Write2NetCDFFile(fn, hw, "hw", ("NPh",), units="eV")
        Write2NetCDFFile(fn, U, "U", ("NPh", "NPh"), units="None")
        Write2NetCDFFile(fn, DynMat, "DynMat", ("NPh", "NPh"), units="eV**2")
        Write2NetCDFFile(fn, SigL.real, "ReSigL", ("NWl", "Nsl", "Nsl"), units="eV**2")
        Write2NetCDFFile(fn, SigL.imag, "ImSigL", ("NWl", "Nsl", "Nsl"), units="eV**2")
        Write2NetCDFFile(fn, SigR.real, "ReSigR", ("NWl", "Nsr", "Nsr"), units="eV**2")
        Write2NetCDFFile(fn, SigR.imag, "ImSigR", ("NWl", "Nsr", "Nsr"), units="eV**2")
        for label, a in (("Friction", Friction), ("NC", NC), ("NCP", NCP), ("zeta1", zeta1), ("zeta2", zeta2)):
            Write2NetCDFFile(fn, a, label, ("NPh", "NPh"), units="eV**2")
    finally:
        fn.close()


def ReadNetCDFVar(file, var):
    """One variable of a file by name (myio.py:238-243)."""
    f = _open(file)
    try:
        return _var(f, var)
    finally:
        f.close()


def ReadMDNCFile(filename):
    """Geometry written by the MD set-up: UnitCell, XYZ, DynamicAtoms, AtomList (myio.py:245-264)."""
    f = _open(filename)
    try:
        md = _Record(filename)
        md.cell, md.xyz = _var(f, "UnitCell"), _var(f, "XYZ")
        md.dynatom, md.atomlist = _var(f, "DynamicAtoms"), _var(f, "AtomList")
    finally:
        f.close()
    return md


def ord2idx(order):
    """1-based atom order -> 0-based degree-of-freedom index list (myio.py:297-303)."""
    order = N.asarray(order, dtype=int)
    return (3 * (order[:, None] - 1) + N.arange(3)[None, :]).ravel()


def ReadDynmat(filename, order=None):
    """hw, U of PHrun's Dev*.nc -> real-space dynamical matrix U^T diag(hw^2) U, symmetrised, with the
    mode columns optionally re-ordered to a new atom order (myio.py:266-295).  Returns (dyn, U, hw)."""
    f = _open(filename)
    try:
        hw, U = _var(f, "hw"), _var(f, "U")
    finally:
        f.close()
    if order is not None:
        if 3 * len(order) != len(hw):
            raise ValueError("ReadDynmat: length of order error!")
        U = U[:, ord2idx(order)]
    dyn = (U.T * hw ** 2) @ U
    return 0.5 * (dyn + dyn.T), U, hw


def ReadSig(filename):
    """Lead self-energies on their frequency grid (myio.py:305-319)."""
    f = _open(filename)
    try:
        eph = _Record()
        eph.wl, eph.SigL, eph.SigR = _var(f, "Wlist"), _cplx(f, "SigL"), _cplx(f, "SigR")
    finally:
        f.close()
    return eph


def ReadwbLambda(filename, order=None):
    """Wide-band electronic couplings: (bias, eta, xim, xip, zeta1, zeta2) with bias = muL - muR
    (myio.py:323-339)."""
    f = _open(filename)
    try:
        mus = _var(f, "muLR")
        out = (mus[0] - mus[1],) + tuple(_var(f, k) for k in ("eta", "xim", "xip", "zeta1", "zeta2"))
    finally:
        f.close()
    return out


def ReadLambda(filename, w0, order=None):
    """Energy-resolved couplings at the grid point nearest to w0 (myio.py:342-369):
    eta = -sym(ImPir2)/w, zeta2 = -asym(ImPir2)/(w bias), xim = -asym(RePir2)/bias,
    zeta1 = sym(RePir2)/bias, xip = -pi sym(ReLamLR)/w, with sym/asym = (A +- A^T)/2.
    Returns (bias, eta, xim, xip, zeta1, zeta2)."""
    from .functions import nearest
    f = _open(filename)
    try:
        wl, mus = _var(f, "wl"), _var(f, "muLR")
        i = nearest(w0, wl)

        def at(name):
            if name not in f.variables:
                raise KeyError("%s: no variable '%s'" % (filename, name))
            return N.array(f.variables[name][i], dtype=float)
        im2, re2, lam = at("ImPir2"), at("RePir2"), at("ReLamLR")
    finally:
        f.close()
    bias, w00 = mus[0] - mus[1], wl[i]
    eta = -(im2 + im2.T) / 2 / w00
    zeta2 = -(im2 - im2.T) / 2 / w00 / bias
    xim = -(re2 - re2.T) / 2 / bias
    zeta1 = (re2 + re2.T) / 2 / bias
    xip = -N.pi * (lam + lam.T) / 2 / w00
    return bias, eta, xim, xip, zeta1, zeta2
