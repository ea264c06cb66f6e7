import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, bench
from pymd_b200 import noise as PN, _lib
Be = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nmd = int(sys.argv[2]) if len(sys.argv) > 2 else 2 ** 13
me = bench.build(Be, nmd, 0, False, 7)
b = me.baths[0]
fac = b.spectral_factors()
nf, nc, ldb = nmd // 2 + 1, b.nc, b.ldb
def T(f, name, n=2):
    for i in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%-28s %.4f s" % (name, dt)); return r
lre = T(lambda: torch.as_tensor(np.ascontiguousarray(fac.L.real)).cuda(), "H2D L.real (%.0f MB)" % (fac.L.real.nbytes / 1e6))
lim = T(lambda: torch.as_tensor(np.ascontiguousarray(fac.L.imag)).cuda(), "H2D L.imag")
xi = T(lambda: PN.white_noise(nf, nc, ldb, 1, 1), "philox normals")
out = torch.empty((nmd, nc, ldb), dtype=torch.float64, device="cuda")
per_col = _lib.load().pymd_noise_work_bytes(nmd, nc, 1)
cw = max(32, min(ldb, (PN.MAX_WORK_BYTES // per_col) // 32 * 32))
work = torch.empty(per_col * cw, dtype=torch.uint8, device="cuda")
def gen():
    for n0 in range(0, ldb, cw):
        _lib.call("pymd_noise_generate", lre.data_ptr(), lim.data_ptr(), xi.data_ptr(), nmd, nc, ldb, n0, min(ldb, n0 + cw),
                  1.0, out.data_ptr(), work.data_ptr(), _lib.stream_ptr())
T(gen, "shape+fft (%d trajectory chunks)" % ((ldb + cw - 1) // cw))
T(lambda: b.gnoi(), "bath.gnoi() total")
