"""times ensemble_powerspec on a [nmd][nd][B] record: python tools/spec_case.py [nmd nd B]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymd_b200.spectrum import ensemble_powerspec
nmd, nd, B = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (32768, 60, 1024)))
x = torch.randn((nmd, nd, B), dtype=torch.float64, device="cuda")
for label, env in (("fused", None), ("separate", "0"), ("fused", None)):
    if env is None: os.environ.pop("PYMD_POWER_FUSED", None)
    else: os.environ["PYMD_POWER_FUSED"] = env
    ensemble_powerspec(x, 1.0, nmd, B, False, host=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): s = ensemble_powerspec(x, 1.0, nmd, B, False, host=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%-9s %.2f ms  %.0f GB/s of the record read once  sum %.6e" % (label, ms, x.numel() * 8 / ms / 1e6, float(s.sum())))
