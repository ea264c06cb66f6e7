"""times the noise generation of the cfg2 baths (Philox + shaping + inverse FFT): python tools/noise_case.py [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pymd_b200.noise as _n
if os.environ.get("NOISE_WORK_GB"):
    _n.MAX_WORK_BYTES = int(os.environ["NOISE_WORK_GB"]) << 30
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
c = bench.workload("cfg2", 1)
m = bench.build(c, B, 0, True, 7)
m.enable_record_overlay()
for b in m.baths:
    b.spectral_factors()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    per = []
    for b in m.baths:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b._gnoi(segment=rep); e1.record(); per.append((e0, e1))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("noise generation %.1f ms:" % (dt * 1e3), " ".join("nc=%d %.1f ms" % (b.nc, a.elapsed_time(z)) for b, (a, z) in zip(m.baths, per)))
