"""Wall-clock breakdown of one e2e sub-ensemble segment (cfg2): where does Run() spend its time?"""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bench
Be = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nmd = int(sys.argv[2]) if len(sys.argv) > 2 else 2 ** 15
me = bench.build(Be, nmd, 0, True, 7)
t = {}
def tick(name, t0):
    torch.cuda.synchronize(); t[name] = t.get(name, 0.0) + time.perf_counter() - t0
t0 = time.perf_counter()
for b in me.baths: b.spectral_factors()
tick("host_factors(setup)", t0)
for rep in range(2):
    t = {k: v for k, v in t.items() if k.startswith("host")}
    t0 = time.perf_counter(); me.initialise(); me.ResetHis(); tick("init", t0)
    for i, b in enumerate(me.baths):
        t0 = time.perf_counter(); b.gnoi(); tick("gnoi_bath%d" % i, t0)
    t0 = time.perf_counter(); me.ResetSavepq(True); tick("reset_savepq", t0)
    t0 = time.perf_counter(); me.steps(nmd); tick("steps", t0)
    t0 = time.perf_counter(); me.GetPower(); tick("getpower", t0)
    t0 = time.perf_counter(); c = [float(b._cur_dev[:, :Be].mean().item()) for b in me.baths]; tick("currents", t0)
    print(json.dumps({k: round(v, 4) for k, v in t.items()}))
print("traj-steps/s in steps():", Be * nmd / t["steps"])
