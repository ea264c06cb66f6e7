#!/usr/bin/env python3
"""Device vs host time of the set-up tables (csrc/setup.cu): the friction kernel of the
config-4 stretch shape (ml=2000, nw=200, nc=81) and the harmonic spectra of a 96-dof junction
at 2048 frequencies.  Prints one JSON line per table (CUDA events; host = numpy on this box)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pymd_b200 import _lib, harmonic as H, synth  # noqa: E402
from pymd_b200.functions import flinterp  # noqa: E402
from pymd_b200.phbath import gamt  # noqa: E402


def dev_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda", 0)
    j = synth.synth(510, 81, 81, seed=4)
    gam = -np.imag(j.SigL[1:]) / j.wl[1:, None, None]
    ml, nw = 2000, 200
    wl = np.array([0.06 * i / nw for i in range(nw)])
    tl = np.array([4.0 * i for i in range(ml)])
    t0 = time.perf_counter()
    ref = gamt(tl, wl, j.wl[1:], gam)
    host_s = time.perf_counter() - t0
    g = flinterp(wl, j.wl[1:], gam).reshape(nw, -1)
    ne = g.shape[1]
    tl_d, w_d, g_d = (torch.as_tensor(x, device=dev) for x in (tl, wl, np.ascontiguousarray(g)))
    out = torch.empty((ml, ne), dtype=torch.float64, device=dev)
    sp = _lib.stream_ptr()
    ms = dev_time(lambda: _lib.call("pymd_gamt", tl_d.data_ptr(), ml, w_d.data_ptr(), nw, g_d.data_ptr(), ne, 0.0,
                                    wl[-1] / np.pi / nw, out.data_ptr(), sp))
    err = np.abs(out.cpu().numpy().reshape(ref.shape) - ref).max() / np.abs(ref).max()
    flops = 2.0 * ml * nw * ne
    print(json.dumps({"table": "gamt", "ml": ml, "nw": nw, "nc": 81, "device_ms": ms, "host_ms": host_s * 1e3,
                      "device_tflops": flops / ms / 1e9, "max_rel_err": err}))

    j = synth.synth(96, 27, 27, seed=9)
    nwh = 2048
    wl = np.linspace(1e-4, 0.025, nwh)
    se = np.zeros((nwh, 96, 96), dtype=complex)
    cl = np.arange(27)
    se[:, cl[:, None], cl[None, :]] = flinterp(wl, j.wl, j.SigL)
    se[:, (69 + cl)[:, None], (69 + cl)[None, :]] = flinterp(wl, j.wl, j.SigR)
    noise = -2.0 * se.imag + 0j
    t0 = time.perf_counter()
    Ds = H.Drs(wl, j.DynMat, se)
    rd, ru = H.DDos(wl, Ds), H.UU(wl, Ds, noise)
    host_s = time.perf_counter() - t0
    wl_d, dyn_d, se_d, pi_d = (torch.as_tensor(x, device=dev) for x in (wl, j.DynMat, se, noise))
    dd, uu = (torch.empty(nwh, dtype=torch.float64, device=dev) for _ in range(2))
    ms = dev_time(lambda: _lib.call("pymd_harmonic", wl_d.data_ptr(), nwh, dyn_d.data_ptr(), se_d.data_ptr(),
                                    pi_d.data_ptr(), 96, 1e-10, dd.data_ptr(), uu.data_ptr(), sp))
    err = max(np.abs(dd.cpu().numpy() - rd).max() / np.abs(rd).max(), np.abs(uu.cpu().numpy() - ru).max() / np.abs(ru).max())
    print(json.dumps({"table": "harmonic", "n": 96, "nw": nwh, "device_ms": ms, "host_ms": host_s * 1e3,
                      "max_rel_err": err}))

    # noise covariances: config 2 (electron bath nc=60 complex, phonon lead nc=15 real; nmd=2^15) and the
    # config-4 stretch lead (nc=81 real, nmd=2^13)
    from pymd_b200 import noise as PN
    j = synth.synth(60, 15, 15, seed=2)
    gam = -np.imag(j.SigL[1:]) / j.wl[1:, None, None]
    j3 = synth.synth(96, 27, 27, seed=9)
    j4 = synth.synth(510, 81, 81, seed=4)
    gam4 = -np.imag(j4.SigL[1:]) / j4.wl[1:, None, None]
    cases = [("ebath nc=60 complex", lambda d: PN.electron_factors(j.efric, j.exim, j.exip, 0.5, 300.0, 1.0, 1.0, 32768, device=d)),
             ("phbath nc=15 real", lambda d: PN.phonon_factors(gam, j.wl[1:], 300.0, 0.03, 1.0, 32768, device=d)),
             ("ebath nc=96 complex (config 3)", lambda d: PN.electron_factors(j3.efric, j3.exim, j3.exip, 0.5, 300.0, 1.0, 2.0, 16384, device=d)),
             ("phbath nc=81 real", lambda d: PN.phonon_factors(gam4, j4.wl[1:], 300.0, 0.03, 15.0, 8192, device=d))]
    for name, fn in cases:
        t0 = time.perf_counter()
        host = fn(None)
        host.device_factors(dev)
        torch.cuda.synchronize()
        host_s = time.perf_counter() - t0
        fn(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d = fn(dev)
        torch.cuda.synchronize()
        dev_s = time.perf_counter() - t0
        hl = host.L[::257] @ np.conj(np.transpose(host.L[::257], (0, 2, 1)))
        dl = d.L[::257] @ np.conj(np.transpose(d.L[::257], (0, 2, 1)))
        print(json.dumps({"table": "spectral factors, " + name, "nf": host.L.shape[0], "device_total_ms": dev_s * 1e3,
                          "host_total_ms": host_s * 1e3, "max_sweeps": int(d.sweeps.max()),
                          "max_err_LLh": float(np.abs(dl - hl).max() / np.abs(hl).max())}))


if __name__ == "__main__":
    main()
