"""NetCDF input readers (pymd_b200/myio.py, reference myio.py:146-369): the EPH.nc schema round
trip, PHrun's Dev*.nc -> dynamical matrix with atom re-ordering, the coupling files, and a junction
set up from a file exactly as examples/AuC2.py does (ReadNewEPHNCFile -> md / phbath / ebath)."""
import numpy as np
import pytest
from scipy.io import netcdf_file

from pymd_b200 import myio, synth


@pytest.fixture()
def eph_file(tmp_path):
    j = synth.synth(12, 3, 6, seed=2, nc_e=12)
    path = str(tmp_path / "EPH.nc")
    hw2, U = np.linalg.eigh(j.DynMat)
    myio.WriteEPHNCfile(path, j.wl, np.sqrt(np.abs(hw2)), U.T, j.DynMat, j.SigL, j.SigR, j.efric, j.exim, j.exip,
                        j.zeta1, j.zeta2)
    return path, j


def test_eph_round_trip(eph_file):
    path, j = eph_file
    e = myio.ReadNewEPHNCFile(path)
    assert e.filename == path
    for got, want in ((e.wl, j.wl), (e.DynMat, j.DynMat), (e.SigL, j.SigL), (e.SigR, j.SigR), (e.efric, j.efric),
                      (e.xim, j.exim), (e.xip, j.exip), (e.zeta1, j.zeta1), (e.zeta2, j.zeta2)):
        assert np.array_equal(got, want)
    assert e.SigL.shape[1] == 3 and e.SigR.shape[1] == 6 and np.iscomplexobj(e.SigL)      # unequal leads survive
    old = myio.ReadEPHNCFile(path)
    assert not hasattr(old, "zeta1") and np.array_equal(old.xim, j.exim)
    s = myio.ReadSig(path)
    assert np.array_equal(s.SigR, j.SigR) and np.array_equal(s.wl, j.wl)
    assert np.array_equal(myio.ReadNetCDFVar(path, "Friction"), j.efric)
    with pytest.raises(KeyError):
        myio.ReadNetCDFVar(path, "nope")


def test_read_dynmat_reorders_atoms(eph_file):
    path, j = eph_file
    dyn, U, hw = myio.ReadDynmat(path)
    assert np.allclose(dyn, j.DynMat, atol=1e-15) and np.array_equal(dyn, dyn.T)
    order = [3, 1, 4, 2]
    assert myio.ord2idx(order).tolist() == [6, 7, 8, 0, 1, 2, 9, 10, 11, 3, 4, 5]
    dyn2, U2, _ = myio.ReadDynmat(path, order=order)
    idx = myio.ord2idx(order)
    assert np.allclose(dyn2, j.DynMat[np.ix_(idx, idx)], atol=1e-15)
    with pytest.raises(ValueError):
        myio.ReadDynmat(path, order=[1, 2])


def test_coupling_and_geometry_files(tmp_path):
    rng = np.random.default_rng(0)
    n, nw = 6, 5
    f = netcdf_file(str(tmp_path / "Lambda.nc"), "w")
    f.createDimension("two", 2); f.createDimension("nw", nw); f.createDimension("n", n)
    v = f.createVariable("muLR", "d", ("two",)); v[:] = [0.3, -0.1]
    v = f.createVariable("wl", "d", ("nw",)); v[:] = np.linspace(0.01, 0.05, nw)
    mats = {}
    for name in ("ImPir2", "RePir2", "ReLamLR"):
        mats[name] = rng.standard_normal((nw, n, n))
        v = f.createVariable(name, "d", ("nw", "n", "n")); v[:] = mats[name]
    for name in ("eta", "xim", "xip", "zeta1", "zeta2"):
        mats[name] = rng.standard_normal((n, n))
        v = f.createVariable(name, "d", ("n", "n")); v[:] = mats[name]
    f.close()
    bias, eta, xim, xip, z1, z2 = myio.ReadwbLambda(str(tmp_path / "Lambda.nc"))
    assert abs(bias - 0.4) < 1e-15 and np.array_equal(eta, mats["eta"]) and np.array_equal(z2, mats["zeta2"])
    bias, eta, xim, xip, z1, z2 = myio.ReadLambda(str(tmp_path / "Lambda.nc"), 0.031)
    i, w = 2, 0.03                                        # nearest grid point
    a, b, c = mats["ImPir2"][i], mats["RePir2"][i], mats["ReLamLR"][i]
    assert np.allclose(eta, -(a + a.T) / 2 / w) and np.allclose(z2, -(a - a.T) / 2 / w / bias)
    assert np.allclose(xim, -(b - b.T) / 2 / bias) and np.allclose(z1, (b + b.T) / 2 / bias)
    assert np.allclose(xip, -np.pi * (c + c.T) / 2 / w)
    assert np.allclose(eta, eta.T) and np.allclose(xim, -xim.T)      # the symmetries ebath.CheckEmat wants

    g = netcdf_file(str(tmp_path / "MD.nc"), "w")
    g.createDimension("three", 3); g.createDimension("na", 4)
    for name, dims, val in (("UnitCell", ("three", "three"), np.eye(3) * 7.0), ("XYZ", ("na", "three"), rng.random((4, 3))),
                            ("DynamicAtoms", ("na",), [1, 2, 3, 4]), ("AtomList", ("na",), [79, 6, 6, 79])):
        v = g.createVariable(name, "d", dims); v[:] = val
    g.close()
    m = myio.ReadMDNCFile(str(tmp_path / "MD.nc"))
    assert m.cell.shape == (3, 3) and m.xyz.shape == (4, 3) and m.atomlist.tolist() == [79, 6, 6, 79]


def test_not_netcdf3_is_reported(tmp_path):
    p = tmp_path / "x.nc"
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises((IOError, OSError, TypeError, ValueError)):
        myio.ReadEPHNCFile(str(p))


def test_junction_from_file_drives_the_bath_setup(eph_file):
    """examples/AuC2.py:120-171 pattern: the file's matrices feed phbath.gmem and ebath's checks."""
    path, j = eph_file
    from pymd_b200.ebath import ebath
    from pymd_b200.phbath import phbath
    e = myio.ReadNewEPHNCFile(path)
    pl = phbath(300.0, [0], 0.03, 40, 2.0, 64, ml=8, sig=e.SigL, gwl=e.wl)
    pl.gmem()
    ref = phbath(300.0, [0], 0.03, 40, 2.0, 64, ml=8, sig=j.SigL, gwl=j.wl)
    ref.gmem()
    assert np.array_equal(np.asarray(pl.kernel), np.asarray(ref.kernel)) and pl.nc == 3
    eb = ebath(list(range(4)), 300.0, 2.0, 64, wmax=1.0, nw=50, bias=0.2, efric=e.efric, exim=e.xim, exip=e.xip,
               zeta1=e.zeta1, zeta2=e.zeta2)
    assert eb.nc == 12 and eb.ml == 1
