#!/usr/bin/env python3
"""Benchmark of the generalized-Langevin stepping path (BASELINE.json metric: trajectory-steps/s).

  python bench.py --gpus N --steps K --warmup W                 # our CUDA path, one rank per GPU
  python bench.py --impl reference --gpus N --steps K ...       # the reference's CPU path on host cores
  python bench.py --config cfg3|cfg4|cfg4s|sweep:<nd>,<ml>,<B>[,<nc>]   # the other BASELINE.json configurations

Default workload (configs[1] of BASELINE.json, "cfg2"): AuBZ-like junction, nd = 60 system dofs, electron
bath on all 60 dofs at finite bias (all five coupling matrices), left/right phonon baths on 15 dofs each with a
100-step memory kernel, one constraint vector, dt = 1, nmd = 2^15, an ensemble of 4096 trajectories per GPU
(weak scaling).  Synthetic junction (pymd_b200/synth.py): the reference's inputs (EPH.nc, ...) do not ship.

One bench "step" = `md_steps_per_step` consecutive md.vv steps for the whole resident ensemble, chosen so that a
step takes ~0.15 s (8192 md steps for cfg2): with the driver's --steps 20 the timed region is seconds long and
the clock sampler sees sustained FP64 load.  `value` = trajectory-steps/s with the coloured noise already
resident in HBM; `e2e` = the same metric through the md/phbath/ebath class API from host inputs: upload of the
spectral factors, noise generation, nmd steps with trajectories saved, power spectra, device->host read of
spectra and currents (+ NCCL all-reduce), with a breakdown of where the time goes.  The e2e ensemble steps as one
piece when its saved trajectories can overlay the consumed noise rows (md.enable_record_overlay: 129 GB for cfg2),
else in sub-ensembles (ensemble.run_pipelined).
"""
import os

for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

import argparse
import json
import subprocess
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_emit = print
CPU_ARM = "one trajectory per host core on a 256-row noise table (the md.vv loop cost depends neither on nmd nor on the ensemble size)"


# ----------------------------------------------------------------------------- workloads
def workload(name, gpus):
    """BASELINE.json configurations as synthetic junctions.  Returns a dict with the
    junction shape, the md parameters and the ensemble layout (weak: `batch` per GPU; strong: `total` split)."""
    base = dict(seed=4, T=300.0, dT=10.0, nw=200, hwcut=0.03, bias=0.5, constraint=True, nce=None, leads=True,
                scaling="weak", nmd_e2e=None)
    if name == "cfg2":
        c = dict(base, name="cfg2", nd=60, ncl=15, ncr=15, nce=60, ml=100, dt=1.0, nmd=2 ** 15, batch=4096,
                 text="cfg2 AuBZ-like junction: nd=60, electron bath nc=60 at bias 0.5 eV (5 matrices), "
                      "2 phonon baths nc=15 ml=100, 1 constraint, dt=1")
    elif name == "cfg3":
        # examples/Alkane.py as BASELINE states it: electron bath only, all 96 dofs, finite bias, T = 0,
        # 8192 trajectories over the GPUs of the box (strong scaling: 1024 per GPU on 8)
        c = dict(base, name="cfg3", nd=96, ncl=3, ncr=3, nce=96, leads=False, ml=1, dt=2.0, nmd=2 ** 14, T=0.0, dT=0.0,
                 seed=12, scaling="strong", total=8192, batch=8192 // max(gpus, 1),
                 text="cfg3 Alkane-like junction: nd=96, electron bath on all 96 dofs at bias 0.5 eV, T=0, 1 constraint, "
                      "dt=2, 8192 trajectories in total")
    elif name == "cfg4":
        c = dict(base, name="cfg4", nd=42, ncl=15, ncr=15, nce=42, ml=100, dt=15.0, nmd=2 ** 13, batch=4096, seed=14,
                 bias=0.3, dT=20.0, constraint=False,
                 text="cfg4 Au6_3x3 junction as shipped: nd=42, electron bath nc=42, 2 phonon baths nc=15 ml=100, dt=15")
    elif name == "cfg4s":
        c = dict(base, name="cfg4s", nd=510, ncl=81, ncr=81, nce=None, ml=2000, dt=8.0, nmd=2 ** 12, batch=1024, seed=16,
                 T=150.0, constraint=False,
                 text="cfg4 stretch (BASELINE: large nd, Nmem=2000): nd=510, 2 phonon baths nc=81 ml=2000, dt=8")
    elif name.startswith("sweep:"):
        f = [int(x) for x in name.split(":")[1].split(",")]
        nd, ml, B = f[0] // 3 * 3, f[1], f[2]
        nc = f[3] if len(f) > 3 else max(3, (nd // 4) // 3 * 3)
        c = dict(base, name=name, nd=nd, ncl=nc, ncr=nc, nce=None, ml=ml, dt=1.0, nmd=2 ** 12, batch=B, seed=9,
                 constraint=False,
                 text="config-5 sweep point: nd=%d, 2 phonon baths nc=%d ml=%d, nmd=4096" % (nd, nc, ml))
    else:
        raise SystemExit("unknown --config %s" % name)
    return c


def algorithmic(c):
    """algorithmic flops and compulsory HBM bytes per trajectory-step of workload c."""
    nd, ml = c["nd"], c["ml"]
    ncs = [c["ncl"], c["ncr"]] if c["leads"] else []
    heads = sum(3 * 2 * n * n for n in ncs)
    taps = sum(2 * (ml - 1) * n * n for n in ncs) if ml > 1 else 0
    if ml == 1 and ncs:
        heads, taps = 0, 0
    local = 0
    if c["nce"]:
        local = 6 * (2 if c["bias"] != 0 else 1) * c["nce"] ** 2
    total = 2 * nd * nd + heads + taps + local
    nbath = len(ncs) + (1 if c["nce"] else 0)
    s_alg = 8 * (sum(ncs) + (c["nce"] or 0) + nbath + 1)
    out = dict(total=total, dyn=2 * nd * nd, heads=heads, taps=taps, local=local, bytes=s_alg)
    if nd > 64 and ml >= 64 and ncs:
        # per-step pipeline with the frequency-domain delay line (csrc/fdl.cu): partitions of L steps; the direct
        # block-Toeplitz product only meets the rows of the current super-block (L/2 + 1 taps on average), the rest
        # is P (L+1)/L complex nc x nc products per step (8 flops per complex multiply-add) plus two 2L-point FFTs
        L = 32
        t = (8.0 * ml) ** 0.5
        while L < 256 and abs(np.log2(2 * L) - np.log2(t)) < abs(np.log2(L) - np.log2(t)):
            L *= 2
        P = (ml - 3) // L + 1
        out["fdl"] = dict(L=L, P=P,
                          near_executed=sum(2 * n * n * (L / 2.0 + 1.0) for n in ncs),
                          far_executed=sum(8.0 * n * n * P * (L + 1) / L + 2 * n * 5.0 * 2 * L * np.log2(2 * L) / L for n in ncs))
    if ncs and ml >= 24:                            # kFarMinMl (csrc/engine.cuh)
        # on-chip path (csrc/fused3.cu + far.cu; only when the run really used both kernels, see run_ours)
        S = 32
        near = (S + 1) / 2.0 + 1.0                     # average number of taps inside the 32-step super-block
        out["tail"] = sum(2 * n * n * (ml - 1 - near) for n in ncs)        # taps older than the super-block: far field
        out["fused"] = total - out["tail"]
        # executed by the far kernel per trajectory-step and window of 97 taps: two real FFTs of length 128 per bath
        # dof (2.5 n log2 n each) and 65 complex nc x nc products (8 flops per complex multiply-add)
        nseg = max(1, (ml - 3 + 96) // 97)
        out["far_executed"] = sum(nseg * (2 * n * 2.5 * 128 * 7 + 65 * 8 * n * n) for n in ncs) / S
    return out


def default_piece(c):
    """md steps per bench step: a power of two sized so that one step takes roughly 0.15 s (estimated from the
    algorithmic flops at 30 TFLOP/s; both arms derive it from the configuration alone): 8192 for cfg2."""
    fl = algorithmic(c)["total"] * c["batch"]
    est = 0.15 * 30e12 / max(fl, 1.0)
    piece = 2 ** int(round(np.log2(max(est, 1.0))))
    return int(min(max(piece, 32), max(c["nmd"], 32)))


def build(c, batch, traj0, savepq, seed, mode=0, nmd=None):
    from pymd_b200 import synth
    from pymd_b200.ebath import ebath
    from pymd_b200.md import md
    from pymd_b200.phbath import phbath
    nmd = nmd or c["nmd"]
    j = synth.synth(c["nd"], c["ncl"], c["ncr"], seed=c["seed"], constraint=c["constraint"], nc_e=c["nce"])
    m = md(c["dt"], nmd, c["T"], axyz=j.axyz, harmonic=True, dyn=j.DynMat, savepq=savepq,
           constr=None if j.constr is None else j.constr.copy(), batch=batch, seed=seed, traj0=traj0, step_mode=mode)
    m.write_files = False
    baths = []
    if c["nce"]:
        baths.append(ebath(j.cats_e, c["T"], c["dt"], nmd, wmax=1.0, nw=500, bias=c["bias"], efric=j.efric, exim=j.exim,
                           exip=j.exip, zeta1=j.zeta1, zeta2=j.zeta2))
    if c["leads"]:
        pl = phbath(c["T"] + c["dT"] / 2, j.cats_l, c["hwcut"], c["nw"], c["dt"], nmd, ml=c["ml"], sig=j.SigL, gwl=j.wl)
        pr = phbath(c["T"] - c["dT"] / 2, j.cats_r, c["hwcut"], c["nw"], c["dt"], nmd, ml=c["ml"], sig=j.SigR, gwl=j.wl)
        pl.gmem(); pr.gmem()
        baths += [pl, pr]
    for b in baths:
        m.AddBath(b)
    return m


def workload_config(a, c, piece):
    nbytes = 8 * sum(x for x in ([c["nce"] or 0] + ([c["ncl"], c["ncr"]] if c["leads"] else []))) * c["batch"]
    return dict(workload=c["text"] + ", nmd=%d" % c["nmd"], config=c["name"], batch_per_gpu=c["batch"],
                md_steps_per_step=piece, nmd=c["nmd"], savepq=False,
                l2="inputs exceed L2: every md step streams %.1f MB of noise rows per GPU (%.1f GB resident), 126 MB L2"
                   % (nbytes / 2 ** 20, nbytes * c["nmd"] / 2 ** 30),
                cpu_arm=CPU_ARM)


# ----------------------------------------------------------------------------- CPU reference arm
def _cpu_modules():
    """The reference itself (oracle/_ref, built from /root/reference by oracle/build_ref.py) when
    it travelled with the snapshot, else the oracle port."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    os.environ["PYMD_REF_QUIET"] = "1"
    if os.path.exists(os.path.join(ref, "md.py")):
        sys.path.insert(0, ref)
        import md as rmd, phbath as rph, ebath as reb   # noqa: E401
        return "reference", rmd.md, rph.phbath, reb.ebath
    from oracle import pymd_oracle as O
    return "port", O.md, O.phbath, O.ebath


def _cpu_system(c, nmd_cpu, seed):
    from pymd_b200 import synth
    kind, MD, PH, EB = _cpu_modules()
    j = synth.synth(c["nd"], c["ncl"], c["ncr"], seed=c["seed"], constraint=c["constraint"], nc_e=c["nce"])
    m = MD(c["dt"], nmd_cpu, c["T"], axyz=j.axyz, harmonic=True, dyn=j.DynMat, savepq=False,
           constr=None if j.constr is None else j.constr.copy())
    baths = []
    if c["nce"]:
        baths.append(EB(j.cats_e, c["T"], c["dt"], nmd_cpu, wmax=1.0, nw=500, bias=c["bias"], efric=j.efric, exim=j.exim,
                        exip=j.exip, zeta1=j.zeta1, zeta2=j.zeta2))
    if c["leads"]:
        pl = PH(c["T"] + c["dT"] / 2, j.cats_l, c["hwcut"], c["nw"], c["dt"], nmd_cpu, ml=c["ml"], sig=j.SigL, gwl=j.wl)
        pr = PH(c["T"] - c["dT"] / 2, j.cats_r, c["hwcut"], c["nw"], c["dt"], nmd_cpu, ml=c["ml"], sig=j.SigR, gwl=j.wl)
        pl.gmem(); pr.gmem()
        baths += [pl, pr]
    for b in baths:
        m.AddBath(b)
    np.random.seed(seed)
    for b in m.baths:
        b.gnoi()
    m.initialise()
    m.ResetHis()
    return kind, m


def _cpu_worker(args):
    """One host core: `warm` untimed then as many timed md.vv steps of ONE trajectory as fit in `budget` seconds
    (at least `nmin`)."""
    c, nmin, warm, budget, nmd_cpu, seed = args
    kind, m = _cpu_system(c, nmd_cpu, seed)
    for _ in range(warm):
        m.vv()
    t0 = time.perf_counter()
    n = 0
    while n < nmin or (time.perf_counter() - t0 < budget and n < 1 << 20):
        m.vv()
        n += 1
    return kind, n, time.perf_counter() - t0


def cpu_throughput(c, nmin, warm, budget, cores=None, nmd_cpu=256):
    """Aggregate trajectory-steps/s of one reference process per host core."""
    import multiprocessing as mp
    cores = cores or min(os.cpu_count() or 1, 64)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(c, nmin, warm, budget, nmd_cpu, 1000 + i) for i in range(cores)])
    kind = res[0][0]
    return dict(value=sum(r[1] / r[2] for r in res), cores=cores, kind=kind, seconds=max(r[2] for r in res),
                steps_per_core=int(np.mean([r[1] for r in res])), per_core=float(np.mean([r[1] / r[2] for r in res])))


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = workload(a.config, a.gpus)
    heavy = c["nd"] > 100 or c["ml"] > 200
    r = cpu_throughput(c, 2 if heavy else 8 * max(a.steps, 1), 1 if heavy else 8 * a.warmup, budget=min(60.0, 1.0 * max(a.steps, 1)))
    sample = ("%d md.vv steps of one trajectory per core (one process per core, %d cores) after warm-up; %s"
              % (r["steps_per_core"], r["cores"], CPU_ARM))
    piece = a.piece or default_piece(c)
    out = dict(impl="reference", metric="trajectory_steps_per_s", value=r["value"], unit="trajectory-steps/s",
               n_gpus=a.gpus, steps=a.steps, warmup=a.warmup, ms_per_step=r["seconds"] * 1e3 / max(a.steps, 1),
               higher_is_better=True, scaling=c["scaling"], vs_baseline=None, dtype="f64", data="synthetic",
               config=workload_config(a, c, piece), gpu_launches=0,
               cpu_baseline=dict(value=r["value"], unit="trajectory-steps/s", cores=r["cores"], kind=r["kind"],
                                 sample=sample, per_core=r["per_core"]),
               e2e=dict(value=r["value"], unit="trajectory-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    _emit(json.dumps(out))


# ----------------------------------------------------------------------------- our arm
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7]
        if not rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        sm = sorted(float(r[0]) for r in rows)
        pw = [float(r[2]) for r in rows if r[2].replace(".", "", 1).isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_min_mhz=sm[0], sm_max_mhz=float(rows[0][1]), reasons=reasons, samples=len(rows),
                    power_w_max=max(pw) if pw else None)


def _peaks():
    p = os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")
    p64 = max(r["tflops"] for r in json.load(open(p))["dmma"]) if os.path.exists(p) else 37.0
    m = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(m)).get("hbm_gbs", 6549.1) if os.path.exists(m) else 6549.1
    return p64, hbm


def shard_check(c, a, rank, world, dev):
    """GPU-side N-GPU invariance: (1) every rank also steps the first 32 trajectories of rank
    (r+1) mod N and compares them BIT FOR BIT with that rank's own result (one all-gather); (2) a fixed global
    ensemble of 8192 trajectories split N ways gives the same averaged currents / kinetic energy / power2 as the
    values computed the same way for any N (they are printed; the N = 1 line is the reference)."""
    import torch
    import torch.distributed as dist
    nmd_s, nst, nsh = 256, 64, 32
    per = c["batch"]

    def run(batch, traj0, steps, savepq):
        m = build(c, batch, traj0, savepq, a.seed + 7, a.mode, nmd=nmd_s)
        m.initialise(stream=0); m.ResetHis()
        for b in m.baths:
            b._gnoi(segment=0)
        if savepq:
            m.ResetSavepq(True)
        m.steps(steps)
        return m

    own = run(per, rank * per, nst, False)
    mine = torch.stack((own._p[:, :nsh], own._q[:, :nsh])).contiguous()
    del own
    # the neighbour's whole shard again on this GPU (same ensemble size, hence the same kernels and tiles)
    nb = run(per, ((rank + 1) % world) * per, nst, False)
    theirs = torch.stack((nb._p[:, :nsh], nb._q[:, :nsh])).contiguous()
    del nb
    res = dict(how="every rank re-runs the shard of rank (r+1) mod N for %d md steps and compares its first %d trajectories "
                   "with that rank's own result (all-gather)" % (nst, nsh))
    if world > 1:
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        ref = gathered[(rank + 1) % world]
    else:
        ref = mine
    bitwise = bool(torch.equal(ref, theirs))
    dev_rel = float((ref - theirs).abs().max() / ref.abs().max())
    flags = torch.tensor([1.0 if bitwise else 0.0, dev_rel], dtype=torch.float64, device=dev)
    if world > 1:
        allf = [torch.empty_like(flags) for _ in range(world)]
        dist.all_gather(allf, flags)
        bitwise = all(bool(f[0].item() == 1.0) for f in allf)
        dev_rel = max(float(f[1].item()) for f in allf)
    res.update(shard_check="bitwise" if bitwise else "rounding", max_rel_dev=dev_rel, trajectories=nsh, md_steps=nst)
    # fixed global ensemble split N ways
    from pymd_b200 import ensemble
    tot = 8192
    loc = tot // world
    g = run(loc, rank * loc, nmd_s, True)
    g.GetPower()
    av = ensemble.ensemble_average(g, tot)
    res["global_8192"] = dict(cur_nW=[float(x) for x in av["cur"]], ekin=av["ekin"],
                              power2_sum=float(av["power2"][:, 1].sum()), power2_max=float(av["power2"][:, 1].max()),
                              note="identical for every N up to the summation order of the all-reduce (compare with the N=1 line)")
    del g
    torch.cuda.empty_cache()
    return res


def run_ours(a):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if local == 0:
        ge.build()                      # no-op when the in-tree library is current
    else:
        for _ in range(600):
            if os.path.exists(ge.LIB) and not ge._stale():
                break
            time.sleep(0.5)
    from pymd_b200 import _lib, ensemble

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    c = workload(a.config, world)
    if a.batch:
        c["batch"] = a.batch
    if a.nmd:
        c["nmd"] = a.nmd
    B, nmd = c["batch"], c["nmd"]
    p64, hbm = _peaks()
    fl = algorithmic(c)

    # ------------------------------------------------------------ resident-ensemble throughput
    t_setup = time.time()
    m = build(c, B, rank * B, False, a.seed, a.mode)
    m.initialise()
    m.ResetHis()
    for b in m.baths:
        b.gnoi()
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup
    # history filled and one bench step sized to ~0.15 s (a power of two of md steps, at most nmd)
    burn = min(max(c["ml"], 32), nmd)
    m.steps(burn)
    torch.cuda.synchronize()
    piece = a.piece or default_piece(c)
    for _ in range(a.warmup):
        m.steps(piece)
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    l0 = _lib.load().pymd_launch_count(m._ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    barrier()
    e0.record()
    for _ in range(a.steps):
        m.steps(piece)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.load().pymd_launch_count(m._ctx) - l0
    clocks = sampler.finish()
    finite = bool(torch.isfinite(m._p).all().item())
    value = B * world * piece * a.steps / (ms * 1e-3)

    # per-kernel device time of the same loop, CUDA events on the launching stream
    import ctypes
    _lib.call("pymd_profile_enable", m._ctx, 1)
    pp = min(piece, 256)
    m.steps(pp)
    ms4 = (ctypes.c_double * 8)()
    cnt4 = (ctypes.c_longlong * 8)()
    _lib.call("pymd_profile_read", m._ctx, ctypes.addressof(ms4), ctypes.addressof(cnt4))
    _lib.call("pymd_profile_enable", m._ctx, 0)
    names = ["tail_gemm", "fused_step", "other_gemm", "stage_kernels", "far_fft"]
    kern = {names[i]: dict(ms=ms4[i], launches=int(cnt4[i])) for i in range(5) if cnt4[i]}
    tot_ms = sum(v["ms"] for v in kern.values())
    # algorithmic flops each kernel class covers (per trajectory-step)
    if "fused_step" in kern and "far_fft" in kern and "fused" in fl:
        cover = dict(fused_step=fl["fused"], far_fft=fl["tail"])
    elif "fused_step" in kern:
        cover = dict(fused_step=fl["total"] - (fl["taps"] if "tail_gemm" in kern else 0), tail_gemm=fl["taps"])
    else:
        cover = dict(tail_gemm=fl["taps"], other_gemm=fl["total"] - fl["taps"])
    per_kernel = {}
    fdl_on = "fdl" in fl and os.environ.get("PYMD_FDL", "1") != "0" and "fused_step" not in kern
    for key, k in kern.items():
        d = dict(share_of_step=k["ms"] / tot_ms, us_per_launch=1e3 * k["ms"] / k["launches"],
                 launches_per_md_step=k["launches"] / pp)
        if fdl_on and key == "tail_gemm":
            # judged on the flops it executes (the taps that meet rows of the current super-block); the whole tail as a
            # direct product is what `whole_step_frac` credits
            ex = fl["fdl"]["near_executed"] * B * pp / (k["ms"] * 1e-3) / 1e12
            d.update(achieved=ex, frac=ex / p64, executed_flops_per_trajectory_step=fl["fdl"]["near_executed"],
                     note="direct taps inside the current %d-step super-block; older rows come from the delay line" % fl["fdl"]["L"])
        elif key in cover and cover[key] > 0:
            ach = cover[key] * B * pp / (k["ms"] * 1e-3) / 1e12
            d.update(achieved=ach, frac=ach / p64, algorithmic_flops_per_trajectory_step=cover[key])
            if key == "far_fft" and "far_executed" in fl:
                # the far field replaces the taps it covers by an FFT convolution: judge the kernel on the
                # flops it executes, and keep the direct-product equivalent beside it
                ex = fl["far_executed"] * B * pp / (k["ms"] * 1e-3) / 1e12
                d.update(achieved=ex, frac=ex / p64, direct_product_equivalent_tflops=ach,
                         note="achieved = executed FFT + per-frequency product flops; the same taps as a direct "
                              "product would be direct_product_equivalent_tflops")
        per_kernel[key] = d
    roof = None
    if per_kernel:
        dom = max(per_kernel, key=lambda n: per_kernel[n]["share_of_step"])
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
        if os.path.exists(tpath) and c["name"] == "cfg2":
            traffic = json.load(open(tpath)).get(dom, {}).get("dram_bytes_per_launch")
        whole = (value / world) * fl["total"] / (p64 * 1e12)
        whole_hbm = (value / world) * fl["bytes"] / (hbm * 1e9)
        roof = dict(bound="tensor", kernel=dom, achieved=per_kernel[dom].get("achieved"), peak=p64, unit="TFLOP/s",
                    frac=per_kernel[dom].get("frac"), traffic=traffic,
                    peak_source="FP64 DMMA.8x8x4 peak measured on this pool (profiles/r01_fp64_peaks.json); "
                                "MEASURED_PEAKS.json has no FP64 entry (cuBLAS DGEMM 8192^3: 35.45 TFLOP/s)",
                    kernels=per_kernel, whole_step_frac=whole, whole_step_frac_of_hbm=whole_hbm,
                    flops_per_trajectory_step=fl)
    del m
    torch.cuda.empty_cache()

    # ------------------------------------------------------------ N-GPU invariance
    shard = None
    if not a.no_shard_check and c["nd"] <= 100:
        shard = shard_check(c, a, rank, world, dev)

    # ------------------------------------------------------------ end to end through the class API
    e2e = None
    value_savepq = None
    if not a.no_e2e:
        nmd_e = c["nmd"]
        Be = min(a.e2e_batch, B)
        while B % Be:
            Be //= 2
        # sub-ensemble size: noise + saved trajectories (+ scratch) must fit; the second noise set of the overlapped
        # pipeline only when it fits as well (run_pipelined checks)
        per_traj = 8 * nmd_e * (sum(_bath_ncs(c)) + 2 * c["nd"])
        free, _ = torch.cuda.mem_get_info(dev)
        while Be > 32 and per_traj * Be + (20 << 30) > free and B % (Be // 2) == 0:
            Be //= 2
        # records overlaid on the consumed noise rows (md.enable_record_overlay): the whole ensemble steps as ONE
        # sub-ensemble with savepq when [ps | qs] + slack fits (129 GB for cfg2) and the two-unit kernel applies
        overlay = False
        me = None
        if not a.no_record_overlay and c["nd"] <= 64 and 2 * c["nd"] >= sum(_bath_ncs(c)) and B % 32 == 0 \
                and 8 * (nmd_e + 36) * 2 * c["nd"] * B + (24 << 30) < free:
            try:
                me = build(c, B, rank * B, True, a.seed + 1, a.mode)
                me.enable_record_overlay()
                overlay, Be = True, B
            except Exception as ex:                     # not applicable (bath layout): sub-ensembles instead
                print("record overlay not used: %s" % ex, file=sys.stderr)
                me = None
                torch.cuda.empty_cache()
        if me is None:
            me = build(c, Be, rank * B, True, a.seed + 1, a.mode)
        for b in me.baths:
            b.spectral_factors()                       # host setup (eigh per frequency), as the reference does once
        h2d = sum(b.spectral_factors().L.nbytes for b in me.baths)
        # untimed warm-up of the same call on ONE sub-ensemble (buffer allocation, streams, first-launch set-up); the
        # timed region below then re-uploads the factors, regenerates the noise and runs every md step of the ensemble
        warm = ensemble.run_pipelined(me, Be, traj0=rank * B, segment=1, overlap=a.overlap, reupload_factors=True)
        pstate = warm["state"]
        del warm
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        try:
            r = ensemble.run_pipelined(me, B, traj0=rank * B, segment=0, overlap=a.overlap, reupload_factors=True, state=pstate)
        except torch.cuda.OutOfMemoryError:
            torch.cuda.empty_cache()
            r = ensemble.run_pipelined(me, B, traj0=rank * B, segment=0, overlap=False, reupload_factors=True)
        t_after = time.perf_counter()
        res = ensemble.allreduce_sums([r["power_sum"], r["power2_sum"], r["cur_sum"], np.array([r["ekin_sum"]])])
        torch.cuda.synchronize()
        barrier()
        t_end = time.perf_counter()
        sec = max_over_ranks(t_end - t0)
        r["seconds"]["after"] = t_end - t_after
        d2h = 2 * nmd_e * 8 + (2 * len(me.baths) + 1) * 8
        br = r["seconds"]
        # HBM roofline of the noise / spectra legs: bytes one pass over the data moves (algorithmic: white noise in,
        # noise rows out; saved trajectories in) against the measured copy bandwidth
        nb_noise = 8 * (nmd_e // 2 + 1 + nmd_e) * sum(_bath_ncs(c)) * B
        nb_spec = 8 * nmd_e * 2 * c["nd"] * B
        e2e = dict(value=B * world * nmd_e / sec, unit="trajectory-steps/s", h2d_bytes_per_step=int(h2d * (B // Be)),
                   d2h_bytes_per_step=int(d2h), seconds=sec, sub_ensembles=B // Be, sub_batch=Be, record_overlay=overlay,
                   step="one Run segment of the whole ensemble in sub-ensembles: factors H2D + noise generation + %d md "
                        "steps with savepq + power spectra + D2H" % nmd_e,
                   breakdown=dict(h2d="inside noise_gen (async copy of %.2f GB per sub-ensemble from pinned memory)" % (h2d / 1e9),
                                  noise_gen=br["noise_gen"], steps=br["steps"], spectra=br["spectra"],
                                  reductions=br["reductions"], d2h=br["d2h"], total=br["total"], setup=br.get("setup"),
                                  after=br.get("after"), overlapped=br["overlap"],
                                  note="device seconds per leg summed over sub-ensembles; with overlap noise_gen of "
                                       "sub-ensemble s+1 runs on a side stream beside the steps of s"),
                   hbm_frac=dict(noise_gen=nb_noise / max(br["noise_gen"], 1e-9) / 1e9 / hbm,
                                 spectra=nb_spec / max(br["spectra"], 1e-9) / 1e9 / hbm,
                                 note="algorithmic bytes (white noise in + noise rows out; saved p, q in) / leg time / "
                                      "%.0f GB/s" % hbm),
                   steps_rate=B * nmd_e / max(br["steps"], 1e-9),
                   mean_current_nW=[float(x) * 243414. / (B * world) for x in res[2]],
                   finite=bool(np.isfinite(res[0]).all() and np.isfinite(res[1]).all()))
        value_savepq = dict(value=e2e["steps_rate"] * world, note="stepping legs of the e2e run alone (savepq=True, "
                            "%d-trajectory sub-ensembles)" % Be)
        del me
        torch.cuda.empty_cache()

    # ------------------------------------------------------------ CPU baseline (rank 0, N == 1)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        heavy = c["nd"] > 100 or c["ml"] > 200
        r = cpu_throughput(c, 2 if heavy else 64, 1 if heavy else 8, budget=a.cpu_seconds)
        cpu = dict(value=r["value"], unit="trajectory-steps/s", cores=r["cores"], kind=r["kind"],
                   per_core=r["per_core"],
                   sample="%d md.vv steps of one trajectory per core (one process per core), %s"
                          % (r["steps_per_core"], CPU_ARM))

    if rank == 0:
        out = dict(metric="trajectory_steps_per_s", value=value, unit="trajectory-steps/s", n_gpus=world,
                   steps=a.steps, warmup=a.warmup, ms_per_step=ms / a.steps, higher_is_better=True, scaling=c["scaling"],
                   vs_baseline=None, dtype="f64", data="synthetic", config=workload_config(a, c, piece), clocks=clocks,
                   gpu_launches=int(launches), e2e=e2e, roofline=roof, cpu_baseline=cpu, value_with_savepq=value_savepq,
                   shard_check=shard, setup_s=t_setup, finite=finite, step_mode=a.mode, timed_seconds=ms * 1e-3)
        _emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _bath_ncs(c):
    return ([c["nce"]] if c["nce"] else []) + ([c["ncl"], c["ncr"]] if c["leads"] else [])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", help="cfg2 (default) | cfg3 | cfg4 | cfg4s | sweep:<nd>,<ml>,<B>[,<nc>]")
    ap.add_argument("--batch", type=int, default=0, help="trajectories per GPU (default: the configuration's)")
    ap.add_argument("--nmd", type=int, default=0)
    ap.add_argument("--piece", type=int, default=0, help="md steps per bench step (default: sized to ~0.15 s)")
    ap.add_argument("--mode", type=int, default=0, help="0 auto, 1 per-step pipeline, 2 on-chip step kernels")
    ap.add_argument("--seed", type=int, default=20240)
    ap.add_argument("--e2e-batch", type=int, default=2048)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-record-overlay", action="store_true", help="e2e: keep saved trajectories and noise in separate arrays (sub-ensembles of --e2e-batch)")
    ap.add_argument("--overlap", action="store_true", help="e2e: generate the next sub-ensemble's noise on a side stream while the current one steps (needs a second set of noise buffers)")
    ap.add_argument("--no-shard-check", action="store_true")
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    # stdout carries exactly ONE JSON line: everything else that libraries print there (e.g. the NCCL
    # version banner) is sent to stderr by pointing fd 1 at fd 2 until the result is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global _emit
    def _emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (line + "\n").encode())
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
