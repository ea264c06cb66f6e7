"""FFT-pass probe: forward transform of n x cols complex columns, timing + a case for ncu."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from pymd_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 21120
x = torch.randn(n, cols, 2, dtype=torch.float64, device="cuda")
work = torch.empty(_lib.load().pymd_fft_work_bytes(n, cols), dtype=torch.uint8, device="cuda")
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _lib.call("pymd_fft_columns", x.data_ptr(), n, cols, -1, work.data_ptr(), _lib.stream_ptr())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
gb = n * cols * 16 / 1e9
print("n=%d cols=%d: %.2f ms, %.2f GB array, %.0f GB/s per read+write of the array" % (n, cols, dt * 1e3, gb, 2 * gb / dt))
