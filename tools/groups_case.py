"""Do independent trajectory groups on separate streams overlap the far-field launches of one group with the step
kernel of another?  python tools/groups_case.py B G   (G ensembles of B/G trajectories, one stream each)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
os.environ.setdefault("PYMD_FUSED_NT", "4")
import torch
import pymd_testlib as T
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
G = int(sys.argv[2]) if len(sys.argv) > 2 else 4
c = dict(T.CASES["aubz"]); c["nmd"] = 8192
ms, streams = [], []
for g in range(G):
    m, j = T.device_md(c, B // G, savepq=False, step_mode=2, seed=1 + g)
    m.initialise(); m.ResetHis()
    for b in m.baths:
        b.gnoi()
    ms.append(m); streams.append(torch.cuda.Stream())
torch.cuda.synchronize()
def run(nsteps, chunk):
    for s0 in range(0, nsteps, chunk):
        for m, st in zip(ms, streams):
            with torch.cuda.stream(st):
                m.steps(chunk)
run(256, 128)
torch.cuda.synchronize()
for chunk in (1024, 256):
    t0 = time.perf_counter()
    run(2048, chunk)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("B=%d groups=%d chunk=%d: %.1f M trajectory-steps/s (%.1f us per 32 steps)" % (B, G, chunk, B * 2048 / dt / 1e6, dt / 64 * 1e6))
