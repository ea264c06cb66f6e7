"""Trajectory-level parity of the CUDA stepping path against the oracle and against the
reference's own golden trajectories, through the md/phbath/ebath class interface.

Tolerance (north star: "stated FP64 relative tolerance"): over <= 128 steps, with identical
white noise, phases and spectral factors, max|dev - oracle| <= 1e-11 * max|oracle| for p, q,
currents and kinetic energy.  The device path evaluates one shared memory-kernel tail instead
of three full convolutions and sums in tensor-core order, so agreement is to rounding, not bit
exact; the dynamics are linear, so the error does not grow exponentially."""
import numpy as np
import pytest

import pymd_testlib as T
from oracle import pymd_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-11


def compare(m, rec, n, tol=TOL):
    B = m.batch
    ps = m.ps if B > 1 else m.ps[..., None]
    qs = m.qs if B > 1 else m.qs[..., None]
    et = m.etot if B > 1 else m.etot[..., None]
    assert T.relerr(ps[..., n], rec["ps"]) < tol
    assert T.relerr(qs[..., n], rec["qs"]) < tol
    assert T.relerr(et[..., n], rec["etot"]) < tol
    for b, bath in enumerate(m.baths):
        cur = bath.cur if B > 1 else bath.cur[:, None]
        scale = max(np.abs(rec["cur"][b]).max(), 1e-300)
        assert np.abs(cur[:, n] - rec["cur"][b]).max() < tol * scale * 10


@pytest.mark.parametrize("case,batch,mode", [("ph2", 1, 1), ("ph2", 3, 1), ("eph", 3, 1), ("chain39", 2, 1),
                                               ("aubz", 2, 1), ("ragged", 33, 1),
                                               ("ph2", 1, 2), ("ph2", 3, 2), ("eph", 3, 2), ("chain39", 2, 2),
                                               ("aubz", 2, 2), ("ragged", 33, 2), ("aubz", 40, 0)])
def test_short_horizon_parity(cuda, case, batch, mode):
    c = T.CASES[case]
    m, j = T.device_md(c, batch, step_mode=mode)
    xi, r = T.draw(c, m, batch, seed=7)
    m.initialise(r=r)
    m.ResetHis()
    for b, bath in enumerate(m.baths):
        bath.gnoi(xi=xi[0][b])
    m.ResetSavepq()
    for _ in range(5):
        m.vv()                              # step-by-step, as tests/test.py drives it
    m.steps(c["nmd"] - 5)                   # and many steps per library call
    assert m.t == c["nmd"]
    for n in sorted(set([0, batch - 1, batch // 2])):
        om, recs = T.run_oracle(c, j, m, xi, r, n, c["nmd"])
        compare(m, recs[0], n)
        pn = m.p if batch > 1 else m.p[..., None]
        qn = m.q if batch > 1 else m.q[..., None]
        assert T.relerr(pn[..., n], om.p) < TOL and T.relerr(qn[..., n], om.q) < TOL
        hist = m.phis if batch > 1 else m.phis[..., None]
        for bath in m.baths:
            if bath.ml > 1:
                assert T.relerr(hist[..., n][:bath.ml][:, bath.cids], om.phis[:bath.ml][:, bath.cids]) < TOL


def test_two_segments_carry_state_and_wrap_noise(cuda):
    """nrep=2 as md.Run does it: t, p, q and the history carry over, fresh noise is drawn per
    segment and indexed t % nmd (phbath.py:196), the last step of a segment reads noise row 0."""
    c = T.CASES["ph2"]
    B = 2
    m, j = T.device_md(c, B)
    xi, r = T.draw(c, m, B, seed=17, nrep=2)
    m.initialise(r=r)
    m.ResetHis()
    recs_dev = []
    for rep in range(2):
        for b, bath in enumerate(m.baths):
            bath.gnoi(xi=xi[rep][b])
        m.ResetSavepq()
        m.steps(c["nmd"])
        recs_dev.append((m.ps.copy(), m.qs.copy(), [bb.cur.copy() for bb in m.baths]))
    assert m.t == 2 * c["nmd"]
    for n in range(B):
        om, recs = T.run_oracle(c, j, m, xi, r, n, c["nmd"], nrep=2)
        for rep in range(2):
            assert T.relerr(recs_dev[rep][0][..., n], recs[rep]["ps"]) < TOL
            assert T.relerr(recs_dev[rep][1][..., n], recs[rep]["qs"]) < TOL
            for b in range(2):
                assert np.abs(recs_dev[rep][2][b][:, n] - recs[rep]["cur"][b]).max() < 10 * TOL * np.abs(recs[rep]["cur"][b]).max()


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name", ["ph2", "eph", "debye"])
def test_against_reference_golden_trajectories(cuda, golden, name, mode):
    """The REFERENCE's own trajectories (golden *.npz): feed its noise and initial state to the
    device path and reproduce ps, qs, currents, kinetic energy and the final history."""
    from pymd_b200 import synth
    from pymd_b200.ebath import ebath
    from pymd_b200.md import md
    from pymd_b200.phbath import phbath
    c, g = golden.cfg(name), golden(name)
    if name == "ph2":
        j = synth.synth(12, 3, 3, seed=11)
        baths = [phbath(c["T"] + c["dT"] / 2, j.cats_l, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=j.SigL, gwl=j.wl),
                 phbath(c["T"] - c["dT"] / 2, j.cats_r, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=j.SigR, gwl=j.wl)]
        constr = None
    elif name == "eph":
        j = synth.synth(12, 3, 3, seed=21, constraint=True)
        baths = [ebath(j.cats_e, c["T"], c["dt"], c["nmd"], wmax=c["ewmax"], nw=c["enw"], bias=c["bias"], efric=j.efric,
                       exim=j.exim, exip=j.exip, zeta1=j.zeta1, zeta2=j.zeta2),
                 phbath(c["T"], j.cats_l, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=j.SigL, gwl=j.wl),
                 phbath(c["T"], j.cats_r, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=j.SigR, gwl=j.wl)]
        constr = j.constr.copy()
    else:
        j = synth.synth(9, 3, 3, seed=31)
        baths = [phbath(c["Tph"], [0, 2], c["debye"], c["nw"], c["dt"], c["nmd"], classical=True, zpmotion=False),
                 ebath(j.cats_e, c["T"], c["dt"], c["nmd"], wmax=c["ewmax"], nw=c["enw"], bias=c["bias"], efric=j.efric,
                       exim=j.exim, exip=j.exip)]
        constr = None
    m = md(c["dt"], c["nmd"], c["T"], axyz=j.axyz, harmonic=True, dyn=j.DynMat, constr=constr, step_mode=mode)
    m.write_files = False
    for b in baths:
        if hasattr(b, "gmem"):
            b.gmem()
        m.AddBath(b)
    m.ResetHis()
    m.p, m.q = g["p0"], g["q0"]
    for rep in range(c["nrep"]):
        for i, b in enumerate(m.baths):
            b.noise = g["noise%d_%d" % (rep, i)]
        m.ResetSavepq()
        m.steps(c["nmd"])
        assert T.relerr(m.ps, g["ps%d" % rep]) < TOL
        assert T.relerr(m.qs, g["qs%d" % rep]) < TOL
        assert T.relerr(m.etot, g["etot%d" % rep]) < TOL
        for i, b in enumerate(m.baths):
            assert np.abs(b.cur - g["cur%d_%d" % (rep, i)]).max() < 10 * TOL * np.abs(g["cur%d_%d" % (rep, i)]).max()
    assert T.relerr(m.p, g["p_end"]) < TOL and T.relerr(m.q, g["q_end"]) < TOL
    for b in m.baths:
        if b.ml > 1:
            assert T.relerr(m.phis[:, b.cids], g["phis_end"][:, b.cids]) < TOL


@pytest.mark.parametrize("case", ["ph2", "eph", "aubz"])
def test_switching_between_step_modes(cuda, case):
    """The per-step pipeline (mode 1) and the fused kernel (mode 2) share all carried state
    (p, q, history ring, tail, cached potential force): alternating between them mid-run, with
    odd step counts that straddle the fused time blocks, still follows the oracle."""
    c = T.CASES[case]
    B = 3
    m, j = T.device_md(c, B)
    xi, r = T.draw(c, m, B, seed=23)
    m.initialise(r=r); m.ResetHis()
    for b, bath in enumerate(m.baths):
        bath.gnoi(xi=xi[0][b])
    m.ResetSavepq()
    plan = [(2, 3), (1, 2), (2, 11), (1, 1), (2, 8), (2, 1), (1, 5)]
    done = 0
    for mode, n in plan:
        m.step_mode = mode
        m.steps(n)
        done += n
    m.step_mode = 2
    m.steps(c["nmd"] - done)
    for n in range(B):
        om, recs = T.run_oracle(c, j, m, xi, r, n, c["nmd"])
        compare(m, recs[0], n)
        assert T.relerr(m.p[..., n], om.p) < TOL and T.relerr(m.q[..., n], om.q) < TOL


@pytest.mark.parametrize("nt,f3", [("1", "1"), ("2", "1"), ("4", "1"), ("2", "0"), ("4", "0")])
@pytest.mark.parametrize("case", ["aubz", "chain39", "ragged", "cfg4"])
def test_fused_tile_sizes(cuda, case, nt, f3, monkeypatch):
    """Every trajectory-tile size of the fused kernel (8, 16, 32 trajectories per CTA), forced
    through the PYMD_FUSED_NT testing knob, against the oracle; 40 trajectories so that tiles are
    partially padded and several CTAs run.  16- and 32-trajectory tiles take the two-unit kernel
    (fused3.cu) by default and the single-unit warp-specialised one (fused2) with PYMD_FUSED3=0."""
    monkeypatch.setenv("PYMD_FUSED_NT", nt)
    monkeypatch.setenv("PYMD_FUSED3", f3)
    c = T.CASES[case]
    B = 40
    m, j = T.device_md(c, B, step_mode=2)
    xi, r = T.draw(c, m, B, seed=31)
    m.initialise(r=r); m.ResetHis()
    for b, bath in enumerate(m.baths):
        bath.gnoi(xi=xi[0][b])
    m.ResetSavepq()
    m.steps(c["nmd"])
    for n in (0, 17, 39):
        om, recs = T.run_oracle(c, j, m, xi, r, n, c["nmd"])
        compare(m, recs[0], n)


def test_fhis_recording(cuda, golden):
    """md.fhis (bath forces of the first evaluation, md.py:343) when asked for."""
    c = T.CASES["ph2"]
    m, j = T.device_md(c, 2, record_fhis=True)
    xi, r = T.draw(c, m, 2, seed=3)
    m.initialise(r=r); m.ResetHis()
    for b, bath in enumerate(m.baths):
        bath.gnoi(xi=xi[0][b])
    m.steps(16)
    om = T.oracle_md(c, j)
    om.initialise(r=r[:, 1]); om.ResetHis()
    for b, ob in enumerate(om.baths):
        ob.gnoi(xi=xi[0][b][:, :, 1], factors=m.baths[b].spectral_factors().as_oracle())
    for _ in range(16):
        om.vv()
    for b in range(2):
        assert T.relerr(m.fhis[b][:16, :, 1], om.fhis[b][:16]) < TOL


def test_history_roundtrip_and_resume(cuda, tmp_path):
    """Checkpoint/resume (md.dump / md.Run resume, md.py:541-582, 654-705): stopping after a
    piece and resuming from the checkpoint gives the same final state as an uninterrupted run."""
    c = dict(T.CASES["ph2"])
    B = 3

    def fresh():
        m, j = T.device_md(c, B, seed=5)
        m.write_files = True
        m.outdir = str(tmp_path)
        m.npie = 4
        return m

    a = fresh()
    a.Run()
    ref_p, ref_q, ref_pow, ref_cur = a.p, a.q, a.power2.copy(), a.curmean.copy()
    import os
    for f in os.listdir(tmp_path):
        os.remove(os.path.join(tmp_path, f))
    # interrupted run: two pieces, then "crash"
    b = fresh()
    b.initialise(); b.ResetHis()
    for bath in b.baths:
        bath.gnoi()
    b.ResetSavepq()
    for i in range(2):
        b.steps(b.nmd // b.npie)
        b.dump(i, 0)
    del b
    r = fresh()
    r.Run()                               # finds MD0.npz with ipie=1, resumes pieces 2,3
    # on resume the memory-kernel tail and the cached force are recomputed from the restored
    # history (different summation order than the carried values): equal to rounding, not bitwise
    assert T.relerr(r.p, ref_p) < 1e-12 and T.relerr(r.q, ref_q) < 1e-12
    assert T.relerr(r.power2, ref_pow) < 1e-12 and T.relerr(r.curmean, ref_cur) < 1e-10
    assert os.path.exists(os.path.join(tmp_path, "kappa.300.0.bath0.run0.dat"))
    assert os.path.exists(os.path.join(tmp_path, "power2.300.0.run0.dat"))


def test_run_matches_oracle_run(cuda):
    """md.Run end to end (initialise, ResetHis, per segment gnoi + nmd steps + GetPower with the
    running mean of md.py:604-607, currents in nW) against the oracle's Run, same draws."""
    c = T.CASES["eph"]
    B = 2
    m, j = T.device_md(c, B)
    m.nrep = 2
    xi, r = T.draw(c, m, B, seed=11, nrep=2)
    # drive Run's pieces by hand to inject the draws
    m.initialise(r=r); m.ResetHis()
    pw = np.zeros((c["nmd"], 2)); pw2 = np.zeros((c["nmd"], 2))
    cm = np.zeros((2, len(m.baths)))
    for rep in range(2):
        for b, bath in enumerate(m.baths):
            bath.gnoi(xi=xi[rep][b])
        m.ResetSavepq(); m.steps(c["nmd"]); m.GetPower()
        pw = (pw * rep + m.power) / (rep + 1.0); pw2 = (pw2 * rep + m.power2) / (rep + 1.0)
        cm[rep] = [bb.cur.mean() * 243414. for bb in m.baths]
    acc = np.zeros(c["nmd"]); acc2 = np.zeros(c["nmd"]); ocm = np.zeros((2, len(m.baths)))
    for n in range(B):
        om = T.oracle_md(c, j)
        om.nrep = 2
        om.Run(r=r[:, n], xi=[[xi[rep][b][:, :, n] for b in range(len(m.baths))] for rep in range(2)],
               factors=[bb.spectral_factors().as_oracle() for bb in m.baths])
        acc += om.power[:, 1] / B; acc2 += om.power2[:, 1] / B; ocm += om.curmean / B
    assert T.relerr(pw[:, 1], acc) < 1e-10 and T.relerr(pw2[:, 1], acc2) < 1e-10
    assert np.allclose(pw[:, 0], om.power[:, 0], rtol=1e-15)
    assert np.abs(cm - ocm).max() < 1e-9 * np.abs(ocm).max()


def test_size_independent_properties_at_scale(cuda, monkeypatch):
    """Properties checked at an ensemble size the oracle cannot follow (B = 4096, AuBZ shape):
    (1) trajectory n is independent of the ensemble it runs in: bit-identical in an ensemble of 4091 and in one
        of 32 placed at the same global indices when both use the same trajectory tile (the kernels differ in
        their summation order, so across tile sizes the agreement is to rounding: 1e-12);
    (2) the dynamics are linear: doubling noise and initial state doubles the trajectory;
    (3) padded trajectories stay exactly at rest; (4) all finite."""
    c = dict(T.CASES["aubz"]); c["nmd"] = 64

    def run(B, traj0=0, scale=1.0):
        m, _ = T.device_md(c, B, savepq=False, seed=9, traj0=traj0)
        m.initialise(); m.ResetHis()
        if scale != 1.0:
            m.p, m.q = scale * np.asarray(m.p), scale * np.asarray(m.q)
        for b in m.baths:
            b.gnoi()
            if scale != 1.0:
                b._noise_dev.mul_(scale)
        m.steps(48)
        return m

    monkeypatch.setenv("PYMD_FUSED_NT", "4")
    big = run(4096 - 5)
    pb, qb = big.p, big.q
    assert np.isfinite(pb).all() and np.isfinite(qb).all()
    assert float(big._p[:, big.batch:].abs().max()) == 0.0
    small = run(32, traj0=64)
    assert np.array_equal(small.p, pb[:, 64:96]) and np.array_equal(small.q, qb[:, 64:96])
    dbl = run(32, traj0=64, scale=2.0)
    assert T.relerr(dbl.p, 2.0 * small.p) < 1e-12
    monkeypatch.delenv("PYMD_FUSED_NT")
    auto = run(32, traj0=64)                    # the tile size the library picks for 32 trajectories
    assert T.relerr(auto.p, pb[:, 64:96]) < 1e-12 and T.relerr(auto.q, qb[:, 64:96]) < 1e-12


def test_external_driver_is_refused(cuda):
    c = T.CASES["ph2"]
    m, _ = T.device_md(c, 1)
    m.AddLMPint(object())
    with pytest.raises(NotImplementedError):
        m.vv()


def test_smoke_case(cuda):
    import smoke_case
    assert smoke_case.run(verbose=False) < 1e-10


@pytest.mark.parametrize("nt", [None, "4"])
@pytest.mark.parametrize("case,nmd,ml", [("chain39", 512, None), ("aubz", 256, None), ("chain39", 512, 250), ("ragged", 256, 120)])
def test_far_field_fft_against_direct_product_and_oracle(cuda, case, nmd, ml, nt, monkeypatch):
    """The FFT far field (csrc/far.cu, super-blocks of 32 steps) against the direct block-Toeplitz
    product (PYMD_FAR=0) over several full 97-entry windows, with ragged steps() calls so that
    super-blocks start at arbitrary times; and against the oracle over the whole horizon.  Memories
    longer than 100 steps (ml = 250, 120) take several 97-row windows per super-block.  nt = "4" forces the
    32-trajectory tile, i.e. the warp-specialised kernel (fused2) where the baths qualify."""
    if nt:
        monkeypatch.setenv("PYMD_FUSED_NT", nt)
    c = dict(T.CASES[case]); c["nmd"] = nmd
    if ml:
        c["ml"] = ml
    batch = 9
    runs = {}
    for far in ("1", "0"):
        monkeypatch.setenv("PYMD_FAR", far)
        m, j = T.device_md(c, batch, step_mode=2)
        xi, r = T.draw(c, m, batch, seed=13)
        m.initialise(r=r)
        m.ResetHis()
        for b, bath in enumerate(m.baths):
            bath.gnoi(xi=xi[0][b])
        m.ResetSavepq()
        done = 0
        for n in [3, 8, 29, 1, 64, 7, 100, 40]:
            n = min(n, nmd - done)
            if n:
                m.steps(n)
                done += n
        m.steps(nmd - done)
        runs[far] = (m, [np.array(b.cur) for b in m.baths], np.array(m.ps), np.array(m.qs))
    (m1, cur1, ps1, qs1), (m0, cur0, ps0, qs0) = runs["1"], runs["0"]
    assert T.relerr(ps1, ps0) < 1e-12 and T.relerr(qs1, qs0) < 1e-12
    for a, b in zip(cur1, cur0):
        assert np.abs(a - b).max() < 1e-11 * max(np.abs(b).max(), 1e-300)
    om, recs = T.run_oracle(c, j, m1, xi, r, 4, nmd)
    compare(m1, recs[0], 4, tol=3e-11)


def test_long_run_statistics_against_the_oracle_ensemble(cuda):
    """North star: long-run kinetic energy and heat currents must agree with the reference path
    within statistical error.  A two-lead junction with a 40 K bias (ph2 shape, quantum baths) is run
    for 1024 steps: 4096 trajectories on the device (Philox noise, random phases) against 48
    independent trajectories of the CPU oracle (numpy RNG).  After a burn-in the time-averaged kinetic
    energy and lead currents of the two ensembles must agree within 4 combined standard errors, and
    the device ensemble must carry a resolved, directed current that roughly balances (the slowest
    modes are still relaxing at this horizon: <J_L> + <J_R> is 1e-7 of 2.4e-5 only after 2048 steps)."""
    c = dict(T.CASES["ph2"]); c.update(nmd=1024, dT=40.0)
    burn = 512
    B = 4096
    m, j = T.device_md(c, B, savepq=False, seed=77)
    m.initialise(); m.ResetHis()
    for b in m.baths:
        b.gnoi()
    m.steps(c["nmd"])
    ek = m._etot[burn:, :B].mean(0).cpu().numpy()                       # per trajectory time averages
    cur = [b._cur_dev[burn:, :B].mean(0).cpu().numpy() for b in m.baths]
    dev = [(x.mean(), x.std(ddof=1) / np.sqrt(B)) for x in [ek] + cur]

    rng = np.random.default_rng(5)
    no = 48
    ora = [[] for _ in range(1 + len(m.baths))]
    nf = c["nmd"] // 2 + 1
    for n in range(no):
        om = T.oracle_md(c, j)
        om.initialise(r=rng.uniform(0, 1, c["nd"]))
        om.ResetHis()
        for b, ob in enumerate(om.baths):
            ob.gnoi(xi=rng.standard_normal((nf, ob.nc)), factors=m.baths[b].spectral_factors().as_oracle())
        om.ResetSavepq()
        for _ in range(c["nmd"]):
            om.vv()
        ora[0].append(om.etot[burn:].mean())
        for b, ob in enumerate(om.baths):
            ora[1 + b].append(ob.cur[burn:].mean())
    for (dm, ds), o in zip(dev, ora):
        o = np.asarray(o)
        om_, os_ = o.mean(), o.std(ddof=1) / np.sqrt(no)
        assert abs(dm - om_) < 4.0 * np.hypot(ds, os_), (dm, ds, om_, os_)
    # what enters from the hot lead leaves into the cold one
    tot = cur[0] + cur[1]
    assert abs(tot.mean()) < 0.35 * abs(cur[0].mean())
    assert cur[0].mean() > 0 > cur[1].mean() and abs(cur[0].mean()) > 5.0 * dev[1][1]


def test_single_trajectory_checkpoint_is_the_reference_netcdf_schema(cuda, tmp_path):
    """batch = 1: md.dump writes MD<j>.nc with the reference's dimensions and variables (md.py:654-705,
    NetCDF-3), and md.Run resumes from it."""
    import os
    from scipy.io import netcdf_file
    c = dict(T.CASES["ph2"])

    def fresh():
        m, j = T.device_md(c, 1, seed=8)
        m.write_files = True
        m.outdir = str(tmp_path)
        m.npie = 4
        return m

    a = fresh()
    a.Run()
    ref_p, ref_pow = a.p, a.power2.copy()
    f = netcdf_file(os.path.join(tmp_path, "MD0.nc"), "r", mmap=False)
    dims = dict(f.dimensions)
    assert dims["nph"] == c["nd"] and dims["two"] == 2 and dims["mem"] == c["ml"] and dims["nmd"] == c["nmd"]
    assert dims["n0"] == 3 and dims["n1"] == 3
    want = {"noise0": ("nnmd", "n0"), "noise1": ("nnmd", "n1"), "ps": ("nnmd", "nph"), "qs": ("nnmd", "nph"),
            "power": ("nnmd", "two"), "power2": ("nnmd", "two"), "energy": ("nnmd",), "p": ("nph",), "q": ("nph",),
            "t": ("one",), "ipie": ("one",), "phis": ("mem", "nph"), "qhis": ("mem", "nph")}
    for k, d in want.items():
        assert f.variables[k].dimensions == d, k
    assert int(f.variables["ipie"][0]) == 3 and int(f.variables["t"][0]) == c["nmd"]
    assert np.array_equal(np.array(f.variables["p"][:]), a.p) and np.array_equal(np.array(f.variables["ps"][:]), a.ps)
    cols = np.concatenate([np.asarray(b.cids) for b in a.baths])
    assert np.abs(np.array(f.variables["phis"][:])[:, cols]).max() > 0    # history of the bath dofs
    f.close()
    for fn in os.listdir(tmp_path):
        os.remove(os.path.join(tmp_path, fn))
    b = fresh()
    b.initialise(); b.ResetHis()
    for bath in b.baths:
        bath.gnoi()
    b.ResetSavepq()
    for i in range(2):
        b.steps(b.nmd // b.npie)
        b.dump(i, 0)
    del b
    r = fresh()
    r.Run()
    assert T.relerr(r.p, ref_p) < 1e-12 and T.relerr(r.power2, ref_pow) < 1e-12


def test_example_driver_script(cuda, tmp_path, monkeypatch):
    """examples/synthetic_junction.py: the reference's script flow (read NetCDF inputs, build EPH.nc, three
    baths, mdrun.Run()) end to end on the device."""
    import importlib.util
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("synthetic_junction", os.path.join(root, "examples", "synthetic_junction.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = str(tmp_path / "run")
    monkeypatch.setattr(sys, "argv", ["synthetic_junction.py", "--batch", "40", "--nmd", "128", "--outdir", out])
    m = mod.main()
    assert m.t == 128 and np.isfinite(m.p).all() and np.isfinite(m.power2).all()
    files = set(os.listdir(out))
    assert {"EPH.nc", "Dev.nc", "MD0.npz", "power2.300.0.run0.dat", "kappa.300.0.bath0.run0.dat"} <= files
    e = mod.ReadNewEPHNCFile(os.path.join(out, "EPH.nc"))
    assert e.DynMat.shape == (48, 48) and np.allclose(e.DynMat, e.DynMat.T)


def test_example_driver_with_device_setup_and_harmonic_check(cuda, tmp_path, monkeypatch):
    """The same script with PYMD_SETUP_DEVICE=1 (friction kernels from pymd_gamt, spectral factors from the
    Jacobi kernel) and --harmonic-check (analytic PP written next to power2 by pymd_harmonic)."""
    import importlib.util
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("synthetic_junction2", os.path.join(root, "examples", "synthetic_junction.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = str(tmp_path / "run")
    monkeypatch.setenv("PYMD_SETUP_DEVICE", "1")
    monkeypatch.setattr(sys, "argv", ["synthetic_junction.py", "--batch", "40", "--nmd", "2048", "--outdir", out,
                                      "--harmonic-check"])
    m = mod.main()
    assert m.t == 2048 and np.isfinite(m.p).all() and np.isfinite(m.power2).all()
    assert all(type(b.spectral_factors()).__name__ == "DeviceSpectralFactors" for b in m.baths)
    tab = np.loadtxt(os.path.join(out, "harmonic.300.0.run0.dat"))
    assert tab.ndim == 2 and tab.shape[1] == 3 and len(tab) == len(m.harmonic_spectra["w"]) > 0
    assert np.isfinite(tab).all() and (tab[:, 1] >= 0).all() and (tab[:, 2] >= 0).all()
    assert np.allclose(tab[:, 1], m.harmonic_spectra["PP"], rtol=1e-6)


def test_thermal_conductance_example_matches_landauer(cuda, tmp_path, monkeypatch):
    """examples/thermal_conductance.py (the reference's thermal.py flow: leads at T +- dT/2, kappa from the
    lead currents, thermal.py:183-201) with the per-temperature set-up on the GPU: for a harmonic junction
    the ensemble conductance must equal the exact stationary current of the linear Langevin equation with
    the baths' discrete memory kernels and generated noise spectra (measured -2.0 %, statistical error
    1.6 %; asserted to 6 %), and lie close to Landauer's formula (measured -6.6 %: the reference's noise
    spectrum and truncated friction kernel are not an exact fluctuation-dissipation pair; asserted to 12 %)."""
    import importlib.util
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("thermal_conductance", os.path.join(root, "examples", "thermal_conductance.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = str(tmp_path / "thermal")
    monkeypatch.setattr(sys, "argv", ["thermal_conductance.py", "--temps", "150", "--batch", "1024", "--outdir", out])
    rows = mod.main()
    T_, kappa, err, kland, kexact = rows[0]
    assert kland > 0 and kappa > 0 and err < 0.03 * kland
    assert abs(kappa / kexact - 1.0) < 0.06
    assert abs(kappa / kland - 1.0) < 0.12 and kexact < kland
    assert os.path.isfile(os.path.join(out, "kappa.150.0.dat"))


@pytest.mark.parametrize("seed", range(8))
def test_randomized_configurations(cuda, seed, monkeypatch):
    """Random junction shapes around the path-selection boundaries (nd up to 63, leads of 3..18 dofs so
    that padded sizes 8/16/24 and the nc <= 16 far-field limit are crossed, memories from below the
    far-field minimum to several 97-row windows, with and without electron bath / constraint, odd ensemble
    sizes, ragged steps() calls) against the oracle on two trajectories."""
    if seed % 2:
        monkeypatch.setenv("PYMD_FUSED_NT", "4")        # 32-trajectory tiles: fused2 where the baths qualify
    rng = np.random.default_rng(1000 + seed)
    nd = 3 * int(rng.integers(5, 22))
    ncl = 3 * int(rng.integers(1, min(7, nd // 6 + 1)))
    ncr = 3 * int(rng.integers(1, min(7, nd // 6 + 1)))
    ml = int(rng.choice([2, 9, 23, 24, 40, 100, 101, 150, 230]))
    eb = bool(rng.integers(0, 2))
    c = dict(nd=nd, ncl=ncl, ncr=ncr, seed=50 + seed, dt=float(rng.choice([1.0, 4.0, 8.0])), nmd=256,
             T=float(rng.choice([10.0, 77.0, 300.0])), dT=float(rng.choice([0.0, 20.0])), ml=ml, nw=60, hwcut=0.03,
             ebath=eb, bias=0.3, constraint=bool(rng.integers(0, 2)), nce=(3 * int(rng.integers(1, nd // 3 + 1)) if eb else None))
    batch = int(rng.choice([1, 2, 7, 33, 70]))
    m, j = T.device_md(c, batch)
    xi, r = T.draw(c, m, batch, seed=seed)
    m.initialise(r=r)
    m.ResetHis()
    for b, bath in enumerate(m.baths):
        bath.gnoi(xi=xi[0][b])
    m.ResetSavepq()
    left = c["nmd"]
    while left:
        n = min(left, int(rng.integers(1, 90)))
        m.steps(n)
        left -= n
    for n in sorted({0, batch - 1}):
        om, recs = T.run_oracle(c, j, m, xi, r, n, c["nmd"])
        compare(m, recs[0], n, tol=3e-11)


@pytest.mark.parametrize("case,batch,nt", [("alkane", 3, "1"), ("alkane", 40, "2"), ("alkane_part", 9, "1"), ("alkane_part", 33, "2")])
def test_local_bath_kernel_for_nd_up_to_96(cuda, case, batch, nt, monkeypatch):
    """csrc/local.cu: one memory-less bath (electron bath at bias with a constraint; at zero bias on a subset
    of the dofs), 64 < nd <= 96, both trajectory tiles, ragged steps() calls, against the oracle and against
    the per-step pipeline."""
    monkeypatch.setenv("PYMD_FUSED_NT", nt)
    c = T.CASES[case]
    res = {}
    for mode in (0, 1):
        m, j = T.device_md(c, batch, step_mode=mode)
        xi, r = T.draw(c, m, batch, seed=3)
        m.initialise(r=r)
        m.ResetHis()
        for b, bath in enumerate(m.baths):
            bath.gnoi(xi=xi[0][b])
        m.ResetSavepq()
        for n in (1, 7, 30, c["nmd"] - 38):
            m.steps(n)
        res[mode] = m
    m = res[0]
    assert T.relerr(m.ps, res[1].ps) < 1e-12 and T.relerr(m.p, res[1].p) < 1e-12
    for n in sorted({0, batch - 1}):
        om, recs = T.run_oracle(c, j, m, xi, r, n, c["nmd"])
        compare(m, recs[0], n)


def _parity_run(case, batch, mode, seed, chunks, trajs, tol=TOL, monkeypatch=None):
    c = T.CASES[case]
    m, j = T.device_md(c, batch, step_mode=mode)
    xi, r = T.draw(c, m, batch, seed=seed)
    m.initialise(r=r)
    m.ResetHis()
    for b, bath in enumerate(m.baths):
        bath.gnoi(xi=xi[0][b])
    m.ResetSavepq()
    done = 0
    for n in chunks:
        n = min(n, c["nmd"] - done)
        if n:
            m.steps(n)
            done += n
    m.steps(c["nmd"] - done)
    for n in trajs:
        om, recs = T.run_oracle(c, j, m, xi, r, n, c["nmd"])
        compare(m, recs[0], n, tol=tol)
        pn = m.p if batch > 1 else m.p[..., None]
        qn = m.q if batch > 1 else m.q[..., None]
        assert T.relerr(pn[..., n], om.p) < tol and T.relerr(qn[..., n], om.q) < tol
        hist = m.phis if batch > 1 else m.phis[..., None]
        for bath in m.baths:
            if bath.ml > 1:
                assert T.relerr(hist[..., n][:bath.ml][:, bath.cids], om.phis[:bath.ml][:, bath.cids]) < tol
    return m


@pytest.mark.parametrize("case,batch,mode", [("cfg3", 2, 0), ("cfg3", 40, 0), ("cfg3", 3, 1),
                                               ("cfg4", 2, 0), ("cfg4", 33, 2), ("cfg4", 3, 1)])
def test_baseline_config_shapes(cuda, case, batch, mode):
    """BASELINE.json configs 3 and 4 in their exact shapes: config 3 = nd 96 (the upper limit of csrc/local.cu),
    electron bath on all 96 dofs at bias 0.5, T = 0, one constraint; config 4 as shipped = nd 42, electron bath
    on 42 dofs, two 15-dof leads with ml = 100, dt = 15.  Auto path selection, the forced on-chip kernels and the
    per-step pipeline against the oracle."""
    _parity_run(case, batch, mode, seed=41, chunks=[1, 9, 31], trajs=sorted({0, batch - 1}))


@pytest.mark.parametrize("case,batch", [("nd129", 2), ("nd129", 34), ("cfg4s", 2), ("nd129e", 3)])
def test_large_junctions_with_memory_baths(cuda, case, batch):
    """nd > 96 with memory baths (the general path; BASELINE configs 4-stretch and 5): nd = 129 with two 30-dof
    leads and ml = 200, a reduced config-4 stretch shape (nd = 150, two 81-dof leads, ml = 300, one constraint)
    and nd = 129 with an electron bath on all dofs at bias plus ragged leads -- 64 steps against the oracle."""
    _parity_run(case, batch, 0, seed=43, chunks=[3, 8, 20], trajs=sorted({0, batch - 1}))


@pytest.mark.parametrize("case,batch,mode", [("con3", 3, 0), ("con3", 3, 1), ("con6", 2, 0), ("con6", 35, 1),
                                               ("con6_local", 9, 0)])
def test_several_constraint_vectors(cuda, case, batch, mode):
    """md.ApplyConstraint with 3 and 6 non-orthogonal, unnormalised vectors (md.py:739-762: every dot product is
    taken with the un-projected vector, so the result is NOT the orthogonal projection; reproduced as is), on a
    junction that would otherwise take the fused kernel, one that would take local.cu, and through the per-step
    pipeline explicitly; also through md.initialise (constraints applied after every mode, md.py:284-285)."""
    m = _parity_run(case, batch, mode, seed=47, chunks=[2, 7, 13], trajs=sorted({0, batch - 1}))
    assert np.asarray(m.constraint).shape[0] == T.CASES[case]["nconstr"]


@pytest.mark.parametrize("case,nmd,ml,batch", [("chain39", 512, 250, 5), ("ragged", 256, 130, 33), ("chain39", 256, 64, 2)])
def test_delay_line_far_field_of_the_per_step_pipeline(cuda, case, nmd, ml, batch, monkeypatch):
    """csrc/fdl.cu: the frequency-domain delay line (partitioned fast convolution, partitions of 32 / 64 steps) that
    the per-step pipeline uses for memories of 64 steps and more, against the direct block-Toeplitz product
    (PYMD_FDL=0) and against the oracle, over several super-blocks, with ragged steps() calls (super-blocks start at
    arbitrary times), a switch to the fused kernel and back in the middle (the delay line is rebuilt from the
    velocity ring), and a history restore."""
    c = dict(T.CASES[case]); c["nmd"] = nmd; c["ml"] = ml
    runs = {}
    for fdl in ("1", "0"):
        monkeypatch.setenv("PYMD_FDL", fdl)
        m, j = T.device_md(c, batch, step_mode=1)
        xi, r = T.draw(c, m, batch, seed=19)
        m.initialise(r=r)
        m.ResetHis()
        for b, bath in enumerate(m.baths):
            bath.gnoi(xi=xi[0][b])
        m.ResetSavepq()
        done = 0
        for i, n in enumerate([5, 40, 1, 70, 33, 64, 9]):
            n = min(n, nmd - done)
            if not n:
                break
            m.step_mode = 2 if i == 3 else 1            # one stretch on the on-chip kernels
            m.steps(n)
            done += n
            if i == 4:
                m.set_history(m.phis)                   # history restore: the delay line starts again from the ring
        m.step_mode = 1
        m.steps(nmd - done)
        runs[fdl] = (m, [np.array(b.cur) for b in m.baths], np.array(m.ps), np.array(m.qs))
    (m1, cur1, ps1, qs1), (m0, cur0, ps0, qs0) = runs["1"], runs["0"]
    assert T.relerr(ps1, ps0) < 1e-12 and T.relerr(qs1, qs0) < 1e-12
    for a, b in zip(cur1, cur0):
        assert np.abs(a - b).max() < 1e-11 * max(np.abs(b).max(), 1e-300)
    n = batch - 1
    om, recs = T.run_oracle(c, j, m1, xi, r, n, nmd)
    compare(m1, recs[0], n, tol=3e-11)
