"""Kernel-level parity (all through the C ABI, on the GPU): tensor-core GEMM, column FFT,
Philox streams, noise pipeline, spectra, initial condition -- each against numpy / the oracle.
FP64 tolerances are stated per test (differences are summation-order only)."""
import numpy as np
import pytest

import pymd_testlib as T
from oracle import pymd_oracle as O

pytestmark = pytest.mark.gpu


def dev_arr(a, cuda):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(cuda)


@pytest.mark.parametrize("M,K,ldb", [(15, 15, 32), (60, 60, 64), (39, 39, 96), (64, 64, 256), (3, 3, 32),
                                       (130, 77, 160), (240, 1488, 4096), (510, 510, 8192)])
def test_dmma_gemm_matches_numpy(cuda, M, K, ldb):
    """Y = alpha A X (+Y): rel. error <= 1e-13 of |A||X| (FP64 DMMA, different summation order)."""
    import torch
    from pymd_b200 import _lib
    rng = np.random.default_rng(M * 1000 + K)
    lda = (K + 31) // 32 * 32
    A = np.zeros((M, lda)); A[:, :K] = rng.standard_normal((M, K))
    X = rng.standard_normal((K, ldb))
    Y0 = rng.standard_normal((M, ldb))
    a, x, y = dev_arr(A, cuda), dev_arr(X, cuda), dev_arr(Y0, cuda)
    _lib.call("pymd_gemm", a.data_ptr(), lda, x.data_ptr(), y.data_ptr(), M, K, ldb, -0.5, 0, _lib.stream_ptr())
    ref = -0.5 * (A[:, :K] @ X)
    scale = (np.abs(A[:, :K]) @ np.abs(X)).max()
    assert np.abs(y.cpu().numpy() - ref).max() <= 1e-13 * scale
    y = dev_arr(Y0, cuda)
    _lib.call("pymd_gemm", a.data_ptr(), lda, x.data_ptr(), y.data_ptr(), M, K, ldb, 2.0, 1, _lib.stream_ptr())
    assert np.abs(y.cpu().numpy() - (Y0 + 2.0 * (A[:, :K] @ X))).max() <= 1e-13 * 4 * scale


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 65536])
@pytest.mark.parametrize("sign", [-1, 1])
def test_column_fft_matches_numpy(cuda, n, sign):
    import torch
    from pymd_b200 import _lib
    rng = np.random.default_rng(n)
    cols = 96
    z = rng.standard_normal((n, cols)) + 1j * rng.standard_normal((n, cols))
    d = torch.as_tensor(z).to(cuda)
    work = torch.empty(_lib.load().pymd_fft_work_bytes(n, cols), dtype=torch.uint8, device=cuda)
    _lib.call("pymd_fft_columns", d.data_ptr(), n, cols, sign, work.data_ptr(), _lib.stream_ptr())
    ref = np.fft.fft(z, axis=0) if sign < 0 else np.fft.ifft(z, axis=0) * n
    assert T.relerr(d.cpu().numpy(), ref) < 5e-15 * max(1, np.log2(n))


@pytest.mark.parametrize("n,cols", [(16384, 4776), (8192, 2400 + 8)])
def test_column_fft_many_tiles_per_cta(cuda, n, cols):
    """Column counts beyond one tile per CTA of the radix-128 passes (74 column CTAs x 32): the cp.async stage ring
    wraps and the last tile is ragged.  Staged and register-prefetch passes do the same arithmetic (bitwise equal);
    sampled columns against numpy."""
    import os
    import torch
    from pymd_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(n + cols)
    z = torch.randn((n, cols, 2), generator=g, dtype=torch.float64)
    outs = {}
    work = torch.empty(_lib.load().pymd_fft_work_bytes(n, cols), dtype=torch.uint8, device=cuda)
    for staged in ("1", "0", "2"):
        os.environ["PYMD_FFT_STAGED"] = staged
        try:
            d = z.to(cuda)
            _lib.call("pymd_fft_columns", d.data_ptr(), n, cols, -1, work.data_ptr(), _lib.stream_ptr())
            outs[staged] = d
        finally:
            del os.environ["PYMD_FFT_STAGED"]
    assert torch.equal(outs["1"], outs["0"]) and torch.equal(outs["2"], outs["0"])
    pick = [0, 31, 32, 2367, 2368, 2369, cols - 33, cols - 1]
    zc = torch.view_as_complex(z)[:, pick].numpy()
    got = torch.view_as_complex(outs["1"])[:, pick].cpu().numpy()
    assert T.relerr(got, np.fft.fft(zc, axis=0)) < 5e-15 * np.log2(n)


def test_myfft_conventions_on_device(cuda):
    """myfft.Fourier1D / iFourier1D (myfft.py:15-52) on device arrays, incl. the round trip the
    reference prints as its self-check (myfft.py:55-76)."""
    import torch
    from pymd_b200.myfft import myfft
    rng = np.random.default_rng(0)
    a = rng.standard_normal((1024, 4))
    f = myfft(0.1, 1024)
    aw = f.Fourier1D(torch.as_tensor(a).to(cuda))
    ref = np.stack([O.Fourier1D(a[:, i], 0.1) for i in range(4)], axis=1)
    assert T.relerr(aw.cpu().numpy(), ref) < 1e-13
    back = f.iFourier1D(aw)
    assert np.abs(back.cpu().numpy() - a).max() < 1e-12
    assert T.relerr(f.iFourier1D(a.astype(complex)), np.stack([O.iFourier1D(a[:, i].astype(complex), 0.1) for i in range(4)], 1)) < 1e-15


def test_philox_streams(cuda):
    """Bit-exact uniforms vs the numpy mirror; normals to 1e-13 (device log/cospi vs numpy);
    shard invariance: trajectory g gets the same numbers whatever traj0/ldb it is generated with."""
    import torch
    from pymd_b200 import _lib
    rows, ldb, seed, sid = 37, 64, 0x1234567890ABCDEF, 0x00030001
    u = torch.empty((rows, ldb), dtype=torch.float64, device=cuda)
    _lib.call("pymd_philox_uniform", u.data_ptr(), rows, ldb, seed, sid, 5, _lib.stream_ptr())
    u1, _ = T.philox_uniforms(rows, ldb, seed, sid, 5)
    assert np.array_equal(u.cpu().numpy(), u1)
    z = torch.empty((rows, ldb), dtype=torch.float64, device=cuda)
    _lib.call("pymd_philox_normal", z.data_ptr(), rows, ldb, seed, sid, 5, _lib.stream_ptr())
    zr = T.philox_normals(rows, ldb, seed, sid, 5)
    assert np.abs(z.cpu().numpy() - zr).max() < 1e-13
    z2 = torch.empty((rows, 32), dtype=torch.float64, device=cuda)
    _lib.call("pymd_philox_normal", z2.data_ptr(), rows, 32, seed, sid, 5 + 20, _lib.stream_ptr())
    assert np.array_equal(z2.cpu().numpy(), z.cpu().numpy()[:, 20:52])
    big = torch.empty((4096, 256), dtype=torch.float64, device=cuda)
    _lib.call("pymd_philox_normal", big.data_ptr(), 4096, 256, 7, 1, 0, _lib.stream_ptr())
    b = big.cpu().numpy()
    assert abs(b.mean()) < 5e-3 and abs(b.std() - 1) < 5e-3 and abs((b ** 4).mean() - 3) < 0.05


@pytest.mark.parametrize("case,batch", [("ph2", 1), ("eph", 3), ("ragged", 40)])
def test_noise_pipeline_matches_oracle(cuda, case, batch):
    """Identical xi and spectral factors -> noise within 1e-12 of max|noise| (FFT algorithm and
    summation order differ from numpy's pocketfft)."""
    c = T.CASES[case]
    m, j = T.device_md(c, batch)
    xi, _ = T.draw(c, m, batch, seed=99)
    om = T.oracle_md(c, j)
    for b, bath in enumerate(m.baths):
        bath.gnoi(xi=xi[0][b])
        got = bath.noise.reshape(c["nmd"], bath.nc, batch)
        fac = bath.spectral_factors().as_oracle()
        for n in range(batch):
            om.baths[b].gnoi(xi=xi[0][b][:, :, n], factors=fac)
            assert T.relerr(got[:, :, n], om.baths[b].noise) < 1e-12, (b, n)


@pytest.mark.parametrize("nmd,batch", [(256, 2), (4096, 2), (16384, 2), (32768, 2), (16384, 1000)])
def test_noise_pipeline_long_records(cuda, nmd, batch):
    """Record lengths whose half-length FFT plans contain the fused radix-128 pass (128; 128 x 16;
    128 x 64; 128 x 128), real-pair loads and real-row stores included; 1000 trajectories: more column tiles than
    CTAs in the radix-128 passes (the cp.async stage ring wraps)."""
    c = dict(T.CASES["ph2"], nmd=nmd)
    m, j = T.device_md(c, batch)
    xi, _ = T.draw(c, m, batch, seed=5)
    om = T.oracle_md(c, j)
    for b, bath in enumerate(m.baths):
        bath.gnoi(xi=xi[0][b])
        got = bath.noise.reshape(nmd, bath.nc, batch)
        fac = bath.spectral_factors().as_oracle()
        for n in sorted({1, batch - 1}):
            om.baths[b].gnoi(xi=xi[0][b][:, :, n], factors=fac)
            assert T.relerr(got[:, :, n], om.baths[b].noise) < 1e-12, (b, n)


def test_noise_with_reference_draw_order(cuda, golden):
    """The reference's own noise (golden, drawn sequentially from the legacy numpy RNG with
    no draw where an eigenvalue is <= 0): replay its draws into xi[f, i] and regenerate on the
    device.  eigh gauge is shared because both sides call the same LAPACK routine on the same
    matrix in this container; on another host only L L^T is defined, hence tolerance 1e-9."""
    from pymd_b200 import synth
    from pymd_b200.phbath import phbath
    c, g = golden.cfg("ph2"), golden("ph2")
    j = synth.synth(12, 3, 3, seed=11)
    np.random.seed(c["seed"])
    np.random.rand(12)                                   # md.initialise consumed one uniform per mode
    for b, (T_, cats, sig) in enumerate(((c["T"] + c["dT"] / 2, j.cats_l, j.SigL), (c["T"] - c["dT"] / 2, j.cats_r, j.SigR))):
        bath = phbath(T_, cats, c["hwcut"], c["nw"], c["dt"], c["nmd"], ml=c["ml"], sig=sig, gwl=j.wl)
        fac = bath.spectral_factors()
        xi = np.zeros((c["nmd"] // 2 + 1, bath.nc))
        for f in range(xi.shape[0]):
            for i in range(bath.nc):
                if fac.lam[f, i] > 0:
                    xi[f, i] = np.random.standard_normal()
        bath.gnoi(xi=xi)
        assert T.relerr(bath.noise, g["noise0_%d" % b]) < 1e-9


@pytest.mark.parametrize("nmd,nd,batch", [(16, 6, 1), (64, 12, 5), (1024, 39, 33), (256, 6, 3), (4096, 6, 3),
                                          (16384, 3, 2), (32768, 3, 2), (32768, 5, 40), (16384, 2, 33), (16384, 40, 70)])
def test_powerspec_matches_oracle(cuda, nmd, nd, batch):
    import torch
    from pymd_b200.spectrum import ensemble_powerspec, powerspec, powerspec2
    rng = np.random.default_rng(nmd)
    ldb = (batch + 31) // 32 * 32
    x = rng.standard_normal((nmd, nd, ldb))
    d = dev_arr(x, cuda)
    dt = 0.7
    s1 = ensemble_powerspec(d, dt, nmd, batch, True)
    s2 = ensemble_powerspec(d, dt, nmd, batch, False)
    r1 = sum(O.powerspec(x[:, :, n], dt, nmd)[:, 1] for n in range(batch))
    r2 = sum(O.powerspec2(x[:, :, n], dt, nmd)[:, 1] for n in range(batch))
    assert T.relerr(s1, r1) < 1e-12 and T.relerr(s2, r2) < 1e-12
    if nmd >= 16384:                                    # the fused last pass against the separate pass + reduction
        import os
        os.environ["PYMD_POWER_FUSED"] = "0"
        try:
            u1 = ensemble_powerspec(d, dt, nmd, batch, True)
        finally:
            del os.environ["PYMD_POWER_FUSED"]
        assert T.relerr(s1, u1) < 1e-13
    one = powerspec(x[:, :, 0], dt, nmd)
    ref = O.powerspec(x[:, :, 0], dt, nmd)
    assert np.allclose(one[:, 0], ref[:, 0], rtol=1e-15) and T.relerr(one[:, 1], ref[:, 1]) < 1e-12
    assert T.relerr(powerspec2(x[:, :, 0], dt, nmd)[:, 1], O.powerspec2(x[:, :, 0], dt, nmd)[:, 1]) < 1e-12


def test_powerspec_chunked_equals_unchunked(cuda):
    import torch
    from pymd_b200 import _lib
    rng = np.random.default_rng(5)
    nmd, nd, ldb = 256, 10, 64
    d = dev_arr(rng.standard_normal((nmd, nd, ldb)), cuda)
    outs = []
    for chunk in (0, 128, 320):
        out = torch.empty(nmd, dtype=torch.float64, device=cuda)
        work = torch.empty(_lib.load().pymd_powerspec_work_bytes(nmd, nd, ldb, chunk), dtype=torch.uint8, device=cuda)
        _lib.call("pymd_powerspec", d.data_ptr(), nmd, nd, ldb, 50, 1.0, 0, out.data_ptr(), work.data_ptr(), chunk, _lib.stream_ptr())
        outs.append(out.cpu().numpy())
    assert T.relerr(outs[1], outs[0]) < 1e-13 and T.relerr(outs[2], outs[0]) < 1e-13


@pytest.mark.parametrize("case", ["ph2", "eph"])
def test_initial_condition_matches_oracle(cuda, case):
    """md.initialise with injected phases, incl. the per-mode constraint projection."""
    c = T.CASES[case]
    B = 7
    m, j = T.device_md(c, B)
    _, r = T.draw(c, m, B, seed=1)
    m.initialise(r=r)
    for n in range(B):
        om = T.oracle_md(c, j)
        om.initialise(r=r[:, n])
        assert T.relerr(m.p[:, n], om.p) < 1e-13 and T.relerr(m.q[:, n], om.q) < 1e-13
    m2, _ = T.device_md(c, 64, seed=42)
    m2.initialise()
    p = m2.p
    assert np.isfinite(p).all() and np.abs(p[:, 0] - p[:, 1]).max() > 0      # distinct trajectories
    m3, _ = T.device_md(c, 32, seed=42, traj0=32)
    m3.initialise()
    assert np.array_equal(m3.p, p[:, 32:64])                             # shard invariance
