"""cuBLAS DGEMM and HBM copy on this box: the FP64 library ceiling next to our own DMMA number."""
import json
import torch

n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
best = 0.0
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = max(best, 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
x = torch.empty(1 << 28, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
bw = 0.0
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
    bw = max(bw, 2 * x.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
print(json.dumps({"cublas_dgemm_8192_tflops": round(best, 2), "copy_gbs": round(bw, 1)}))
