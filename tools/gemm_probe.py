"""Time the tail-shaped tensor-core GEMM (M x K x N) through pymd_gemm for a few shapes."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from pymd_b200 import _lib
_lib.load()
for (M, K, N) in [(120, 1568, 4096), (120, 1568, 16384), (240, 1568, 4096), (60, 64, 4096), (510, 512, 8192), (15, 1568, 4096)]:
    lda = (K + 31) // 32 * 32
    a = torch.randn(M, lda, dtype=torch.float64, device="cuda")
    x = torch.randn(K, N, dtype=torch.float64, device="cuda")
    y = torch.empty(M, N, dtype=torch.float64, device="cuda")
    for _ in range(3):
        _lib.call("pymd_gemm", a.data_ptr(), lda, x.data_ptr(), y.data_ptr(), M, K, N, 1.0, 0, _lib.stream_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _lib.call("pymd_gemm", a.data_ptr(), lda, x.data_ptr(), y.data_ptr(), M, K, N, 1.0, 0, _lib.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(json.dumps(dict(M=M, K=K, N=N, us=round(us, 1), tflops=round(2 * M * K * N / us * 1e-6, 2))))
