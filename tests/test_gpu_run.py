"""md.Run bookkeeping on the device path: checkpoints with recorded bath forces, resumed multi-segment
runs, state-only checkpoints of large ensembles, and operands changed after the first step (bias sweeps on
one md object, re-run gmem(), a new constraint) -- the reference reads them live on every force call
(ebath.py:191-194, phbath.py:197-201, md.py:355-358)."""
import os

import numpy as np
import pytest

import pymd_testlib as T

pytestmark = pytest.mark.gpu


def _fresh(c, B, outdir, **kw):
    m, j = T.device_md(c, B, **kw)
    m.write_files = True
    m.outdir = str(outdir)
    m.npie = 4
    return m, j


def test_dump_with_recorded_bath_forces(cuda, tmp_path):
    """record_fhis=True with write_files=True (the default): the single-trajectory MD<j>.nc holds fhis<i> in
    the reference's (nnmd, nph) layout (md.py:676-678), nc != nph; the ensemble .npz keeps the bath's dofs."""
    from scipy.io import netcdf_file
    c = T.CASES["eph"]
    m, j = _fresh(c, 1, tmp_path, seed=3, record_fhis=True)
    m.Run()
    f = netcdf_file(os.path.join(tmp_path, "MD0.nc"), "r", mmap=False)
    for i, b in enumerate(m.baths):
        v = np.array(f.variables["fhis%d" % i][:])
        assert f.variables["fhis%d" % i].dimensions == ("nnmd", "nph") and v.shape == (c["nmd"], c["nd"])
        assert np.array_equal(v, m.fhis[i])
        other = np.setdiff1d(np.arange(c["nd"]), b.cids)
        assert np.abs(v[:, b.cids]).max() > 0 and (len(other) == 0 or np.abs(v[:, other]).max() == 0)
    f.close()
    d2 = tmp_path / "ens"
    d2.mkdir()
    e, _ = _fresh(c, 3, d2, seed=3, record_fhis=True)
    e.Run()
    z = np.load(os.path.join(d2, "MD0.npz"))
    for i, b in enumerate(e.baths):
        assert z["fhis%d" % i].shape == (c["nmd"], b.nc, 3)
        assert np.array_equal(z["fhis%d" % i], e.fhis[i][:, b.cids])


@pytest.mark.parametrize("light", [False, True])
def test_resumed_two_segment_run_repeats_the_uninterrupted_one(cuda, tmp_path, light):
    """nrep = 2 with an explicit seed: a run interrupted in segment 1 and resumed in a NEW process-like object
    (fresh counters) draws the same noise for every segment as the uninterrupted run (the Philox stream is a
    function of the segment index, not of how many gnoi() calls this process has made).  light: the same with
    state-only end-of-segment checkpoints (checkpoint_bytes_limit = 0)."""
    c = dict(T.CASES["ph2"])
    B = 3

    def make(out):
        m, _ = _fresh(c, B, out, seed=11)
        m.nrep = 2
        if light:
            m.checkpoint_bytes_limit = 0
        return m

    a_dir = tmp_path / "a"; a_dir.mkdir()
    a = make(a_dir)
    a.Run()
    if light:
        z = np.load(os.path.join(a_dir, "MD1.npz"))
        assert "light" in z and "noise0" not in z and "ps" not in z and z["curmean0"].shape == (B,)
    b_dir = tmp_path / "b"; b_dir.mkdir()
    b = make(b_dir)
    b.nrep = 1
    b.Run()                                     # segment 0 complete, then "crash"
    if not light:
        # ... and half of segment 1
        b.nrep = 2
        st = b._resume(1)
        assert st is None
        for bath in b.baths:
            bath._gnoi(segment=1)
        b.ResetSavepq(True)
        for i in range(2):
            b.steps(b.nmd // b.npie)
            b.dump(i, 1)
    del b
    r = make(b_dir)
    r.Run()
    assert T.relerr(r.p, a.p) < 1e-12 and T.relerr(r.q, a.q) < 1e-12
    assert T.relerr(r.power2, a.power2) < 1e-12 and T.relerr(r.power, a.power) < 1e-12
    assert T.relerr(r.curmean[1], a.curmean[1]) < 1e-10
    for x, y in zip(r.baths, a.baths):
        assert T.relerr(x.noise, y.noise) < 1e-13          # segment 1's realisation, not segment 0's again


def test_operands_changed_after_the_first_step_reach_the_device(cuda):
    """ebath.setbias, a changed constraint and a re-run gmem() between steps() calls: the next step uses the
    new operands (as the reference's bforce / ApplyConstraint do) and keeps the velocity history."""
    c = dict(T.CASES["eph"])
    B = 2
    m, j = T.device_md(c, B)
    xi, r = T.draw(c, m, B, seed=5)
    m.initialise(r=r); m.ResetHis()
    for b, bath in enumerate(m.baths):
        bath.gnoi(xi=xi[0][b])
    m.ResetSavepq()
    m.steps(10)
    # oracle twin, driven the same way
    oms = []
    for n in range(B):
        om = T.oracle_md(c, j)
        om.initialise(r=r[:, n]); om.ResetHis()
        for b, ob in enumerate(om.baths):
            ob.gnoi(xi=xi[0][b][:, :, n], factors=m.baths[b].spectral_factors().as_oracle())
        om.ResetSavepq()
        for _ in range(10):
            om.vv()
        oms.append(om)
    # 1. new bias (friction + non-conservative operators change; noise kept, as a script that only calls setbias)
    m.baths[0].setbias(0.7)
    # 2. new constraint vector
    newc = np.zeros((1, c["nd"])); newc[0, 0::3] = 1.0; newc[0, 1] = 0.5
    m.constraint = newc
    # 3. shorter memory, kernel recomputed (the reference keeps the newest ml rows of its history)
    for bath in m.baths[1:]:
        bath.SetMemlen(4)
        bath.gmem()
    m.ml = 4
    m.steps(12)
    for n, om in enumerate(oms):
        om.baths[0].setbias(0.7)
        om.constraint = newc.copy()
        for ob, bath in zip(om.baths[1:], m.baths[1:]):
            ob.ml = 4
            ob.kernel = np.asarray(bath.kernel)
        om.ml = 4
        om.phis, om.qhis = om.phis[:4].copy(), om.qhis[:4].copy()
        for _ in range(12):
            om.vv()
        assert T.relerr(m.p[..., n], om.p) < 1e-11 and T.relerr(m.q[..., n], om.q) < 1e-11
        assert T.relerr(m.ps[10:22, :, n], om.ps[10:22]) < 1e-11


@pytest.mark.parametrize("overlap", [False, True])
def test_pipelined_sub_ensembles_equal_one_run_per_sub_ensemble(cuda, overlap):
    """ensemble.run_pipelined (what bench.py's e2e leg calls): an ensemble of 96 trajectories in three sub-ensembles
    of 32 -- with and without the next sub-ensemble's noise generated on a side stream -- gives exactly the sums of
    three separate runs (initialise, gnoi, steps, GetPower) of the same global trajectory indices."""
    from pymd_b200 import ensemble
    c = dict(T.CASES["eph"], nmd=64, ml=30)
    Be, nsub, seed = 32, 3, 21
    m, _ = T.device_md(c, Be, seed=seed, traj0=5)
    for b in m.baths:
        b.spectral_factors()
    r = ensemble.run_pipelined(m, Be * nsub, traj0=5, segment=0, overlap=overlap, reupload_factors=True)
    assert r["seconds"]["overlap"] == overlap and r["seconds"]["sub_ensembles"] == nsub
    p1 = np.zeros(c["nmd"]); p2 = np.zeros(c["nmd"]); cs = np.zeros(len(m.baths)); cq = np.zeros(len(m.baths)); ek = 0.0
    for s in range(nsub):
        a, _ = T.device_md(c, Be, seed=seed, traj0=5 + s * Be)
        a.initialise(stream=0); a.ResetHis()
        for b in a.baths:
            b._gnoi(segment=0)
        a.ResetSavepq(True)
        a.steps(c["nmd"])
        a.GetPower()
        p1 += a.power_sum; p2 += a.power2_sum
        for i, b in enumerate(a.baths):
            ct = b.cur.mean(0)
            cs[i] += ct.sum(); cq[i] += (ct * ct).sum()
        ek += a.etot.mean(0).sum()
    assert np.array_equal(r["power_sum"], p1) and np.array_equal(r["power2_sum"], p2)
    assert T.relerr(r["cur_sum"], cs) < 1e-12 and T.relerr(r["cur_sq"], cq) < 1e-12 and abs(r["ekin_sum"] / ek - 1) < 1e-12
    assert T.relerr(m.power2[:, 1], p2 / (Be * nsub)) < 1e-15


@pytest.mark.parametrize("case,B", [("aubz", 40), ("chain39", 33)])
def test_overlaid_records_equal_separate_arrays(cuda, case, B, monkeypatch):
    """md.enable_record_overlay: saved trajectories written over the consumed noise rows of one shared buffer (what
    lets 4096 trajectories of config 2 step as one ensemble with savepq).  Same seeds with and without the overlay:
    ps, qs, currents, kinetic energy, final state and the power spectra are bit-identical; the library refuses a
    steps() call that would run past the end of the noise table and kernels that do not support strided records."""
    from pymd_b200 import _lib
    monkeypatch.setenv("PYMD_FUSED_NT", "4")            # the two-unit kernel on 32-trajectory tiles in both runs
    c = dict(T.CASES[case]); c["nmd"] = 128
    out = {}
    for ov in (False, True):
        m, _ = T.device_md(c, B, seed=13)
        if ov:
            m.enable_record_overlay()
            assert m._ps.stride(0) == 2 * c["nd"] * m.ldb and m.baths[0]._noise_dev.stride(0) == sum(b.nc for b in m.baths) * m.ldb
        m.initialise(stream=0); m.ResetHis()
        for b in m.baths:
            b._gnoi(segment=0)
        noise0 = [np.array(b.noise) for b in m.baths]
        m.ResetSavepq(True)
        m.steps(5)
        m.steps(59)
        if ov:
            with pytest.raises(_lib.PymdError):
                m.steps(65)                             # would run past the end of the (overwritten) noise table
        m.steps(64)
        m.GetPower()
        out[ov] = (np.array(m.ps), np.array(m.qs), [np.array(b.cur) for b in m.baths], np.array(m.etot), np.array(m.p),
                   np.array(m.power[:, 1]), np.array(m.power2[:, 1]), noise0)
    a, b = out[False], out[True]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
    for x, y in zip(a[2], b[2]):
        assert np.array_equal(x, y)
    assert np.array_equal(a[5], b[5]) and np.array_equal(a[6], b[6])
    for x, y in zip(a[7], b[7]):
        assert np.array_equal(x, y)                     # the generated noise itself (read before stepping)
