"""Ensemble sharding over the GPUs of one box and the final NCCL average.

Trajectories are independent: the global ensemble [0, B) is cut into contiguous slices, one
per rank; small shared operands (dyn, kernels, spectral factors) are replicated; the white
noise of trajectory g depends only on (seed, g), so results do not depend on the GPU count.
The only collective is one all-reduce(sum) of a few (nmd,) vectors after the time loop.
"""
import numpy as N


def shard(batch_total, rank=None, world=None):
    """(traj0, nlocal) of this rank's contiguous slice."""
    if rank is None or world is None:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
        else:
            rank, world = 0, 1
    base, rem = divmod(int(batch_total), world)
    n = base + (1 if rank < rem else 0)
    t0 = rank * base + min(rank, rem)
    return t0, n


def allreduce_sums(arrays, device=None):
    """Sum a list of host float64 arrays over all ranks with ONE collective (NCCL on GPUs,
    gloo in CPU tests); returns new host arrays.  Identity when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [N.array(a, dtype=N.float64, copy=True) for a in arrays]
    flat = N.concatenate([N.asarray(a, dtype=N.float64).ravel() for a in arrays])
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.as_tensor(flat, dtype=torch.float64).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    flat = t.cpu().numpy()
    out, o = [], 0
    for a in arrays:
        n = int(N.asarray(a).size)
        out.append(flat[o:o + n].reshape(N.shape(a)))
        o += n
    return out


def ensemble_average(mdrun, batch_total):
    """After GetPower()/Run() on every rank: global ensemble means of the spectra, per-bath
    mean currents (nW) and their standard errors, and the mean kinetic energy."""
    from . import units as U
    nb = len(mdrun.baths)
    from .spectrum import time_means
    cur_t = [time_means(b._cur_dev)[:mdrun.batch] for b in mdrun.baths]         # time mean per trajectory
    cur_sum = N.array([float(c.sum().item()) for c in cur_t])
    cur_sq = N.array([float((c * c).sum().item()) for c in cur_t])
    ek = float(time_means(mdrun._etot)[:mdrun.batch].sum().item())
    p1 = getattr(mdrun, "power_sum", N.zeros(mdrun.nmd))
    p2 = getattr(mdrun, "power2_sum", N.zeros(mdrun.nmd))
    p1, p2, cur_sum, cur_sq, ekv = allreduce_sums([p1, p2, cur_sum, cur_sq, N.array([ek])])
    B = float(batch_total)
    mean = cur_sum / B
    var = N.maximum(cur_sq / B - mean ** 2, 0.0)
    dw = 2. * N.pi / mdrun.dt / mdrun.nmd
    w = dw * N.arange(mdrun.nmd)
    return dict(power=N.stack((w, p1 / B), 1), power2=N.stack((w, p2 / B), 1),
                cur=mean * U.curcof, cur_err=N.sqrt(var / max(B - 1, 1)) * U.curcof,
                ekin=float(ekv[0]) / B, nbath=nb)


def run_pipelined(mdrun, batch_total, traj0=0, segment=0, overlap=False, reupload_factors=False, state=None):
    """One md.Run segment (initialise, ResetHis, fresh noise, nmd steps with saved trajectories, GetPower) for
    an ensemble of ``batch_total`` trajectories that is larger than what fits beside its own saved trajectories:
    it is processed in sub-ensembles of ``mdrun.batch`` trajectories (global indices traj0 + s batch ...).  The
    spectral factors of sub-ensemble s+1 are uploaded on a copy stream while s is stepping (``reupload_factors``:
    the inputs cross the bus for every sub-ensemble).  ``overlap=True`` also generates the coloured noise of
    sub-ensemble s+1 on a low-priority side stream, into a second set of noise buffers, while s is stepping: measured
    (profiles/r02_e2e_overlap.txt) the step kernels then slow down by 18 % -- they own 128 SMs with all their
    shared memory, so the noise kernels get the 20 idle SMs plus the gaps between launches -- and the net gain is
    4 %, not worth the second 45 GB; it is off by default and needs the memory to be there.  Every per-trajectory
    result is identical either way: noise and initial phases are functions of (seed, segment, global index).

    ``state``: streams, the second noise set and the accumulators of a previous call (returned as out["state"]); a
    run of several segments passes it on, so that only the first segment allocates.

    Returns dict(power_sum, power2_sum [nmd], cur_sum, cur_sq [nbath], ekin_sum, seconds=dict(...), state): sums over
    the local trajectories, ready for ensemble.allreduce_sums; one device->host copy at the end."""
    import time
    import torch
    from .spectrum import ensemble_powerspec, time_means
    t_enter = time.perf_counter()
    Be = mdrun.batch
    if batch_total % Be:
        raise ValueError("run_pipelined: batch_total must be a multiple of md.batch")
    nsub = batch_total // Be
    dev = mdrun._dev()
    mdrun._ensure()
    baths = mdrun.baths
    nmd, ldb = mdrun.nmd, mdrun.ldb
    if mdrun._ps is None:
        mdrun.ResetSavepq(True)
    if state is not None and state.get("key") == (id(mdrun), Be, nmd, ldb, bool(overlap)):
        overlap, main, side, copy, bufs, acc_p, acc_p2, acc_c, acc_e = state["objs"]
        acc_p.zero_(); acc_p2.zero_(); acc_c.zero_(); acc_e.zero_()
    else:
        want = bool(overlap)
        free, _ = torch.cuda.mem_get_info(dev)
        need = sum(8 * nmd * b.nc * ldb for b in baths)
        overlap = bool(overlap and nsub > 1 and free > need + (30 << 30))      # second noise set + scratch must fit
        main = torch.cuda.Stream(dev, priority=-1)
        side = torch.cuda.Stream(dev, priority=0) if overlap else main
        copy = torch.cuda.Stream(dev, priority=0)             # factor uploads of the next sub-ensemble (copy engine)
        bufs = [[None, None] for _ in baths]
        for i, b in enumerate(baths):
            shape = (nmd, b.nc, ldb)
            cur = b._noise_dev if (b._noise_dev is not None and tuple(b._noise_dev.shape) == shape) else \
                torch.empty(shape, dtype=torch.float64, device=dev)
            bufs[i][0] = cur
            bufs[i][1] = torch.empty(shape, dtype=torch.float64, device=dev) if overlap else cur
        acc_p = torch.zeros(nmd, dtype=torch.float64, device=dev)
        acc_p2 = torch.zeros(nmd, dtype=torch.float64, device=dev)
        acc_c = torch.zeros((2, len(baths)), dtype=torch.float64, device=dev)
        acc_e = torch.zeros(1, dtype=torch.float64, device=dev)
        state = dict(key=(id(mdrun), Be, nmd, ldb, want), objs=(overlap, main, side, copy, bufs, acc_p, acc_p2, acc_c, acc_e))
    staged = {}
    torch.cuda.synchronize(dev)
    ready = [torch.cuda.Event() for _ in range(nsub)]
    done = [torch.cuda.Event() for _ in range(nsub)]
    tev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(nsub)]
    nev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(nsub)]

    def generate(s):
        with torch.cuda.stream(side):
            if s >= 2:
                side.wait_event(done[s - 2])            # the buffers of sub-ensemble s-2 are free again
            nev[s][0].record(side)
            order = list(range(len(baths)))
            evs = {}
            if reupload_factors and s not in staged:
                # this sub-ensemble's inputs cross the bus now: all uploads go to the copy stream at once, smallest
                # first, and the baths are generated in that order, so that the upload of the big electron-bath
                # factors (0.94 GB at config 2) runs beside the noise generation of the leads
                order.sort(key=lambda i: baths[i].spectral_factors().L.nbytes if hasattr(baths[i].spectral_factors(), "_pinned") else 0)
                with torch.cuda.stream(copy):
                    for i in order:
                        fac = baths[i].spectral_factors()
                        fac.drop_device_copy()
                        pair = fac.upload_copy(dev)
                        for t in pair:
                            if t is not None:
                                t.record_stream(side)   # allocated on the copy stream, read by kernels of `side`
                        fac.adopt_device_copy(dev, pair)
                        evs[i] = torch.cuda.Event()
                        evs[i].record(copy)
            for i in order:
                b = baths[i]
                if reupload_factors:
                    if s in staged:
                        side.wait_event(staged[s][1])
                        b.spectral_factors().adopt_device_copy(dev, staged[s][0][i])    # uploaded while the previous sub-ensemble stepped
                    elif i in evs:
                        side.wait_event(evs[i])
                b._gnoi(segment=segment, out=bufs[i][s % 2], traj0=traj0 + s * Be)
            staged.pop(s, None)
            nev[s][1].record(side)
            ready[s].record(side)

    t0 = time.perf_counter()
    generate(0)
    with torch.cuda.stream(main):
        for s in range(nsub):
            if s + 1 < nsub and overlap:
                generate(s + 1)
            main.wait_event(ready[s])
            mdrun.traj0 = traj0 + s * Be
            for i, b in enumerate(baths):
                b._traj0 = mdrun.traj0
                b._noise_dev = bufs[i][s % 2]
                b._noise_version = getattr(b, "_noise_version", 0) + 1
            if reupload_factors and not overlap and s + 1 < nsub:
                with torch.cuda.stream(copy):
                    ups = [b.spectral_factors().upload_copy(dev) for b in baths]
                    for pair in ups:
                        for t in pair:
                            if t is not None:
                                t.record_stream(side)
                    ev = torch.cuda.Event()
                    ev.record(copy)
                staged[s + 1] = (ups, ev)
            tev[s][0].record(main)
            mdrun.initialise(stream=segment)
            mdrun.ResetHis()
            mdrun.steps(nmd)
            tev[s][1].record(main)
            done[s].record(main)
            s1 = ensemble_powerspec(mdrun._qs, mdrun.dt, nmd, Be, wsq=True, host=False)
            s2 = ensemble_powerspec(mdrun._ps, mdrun.dt, nmd, Be, wsq=False, host=False)
            tev[s][2].record(main)
            if s + 1 < nsub and not overlap:
                generate(s + 1)
            acc_p += s1
            acc_p2 += s2
            # time means per trajectory: one pass over each [nmd][ldb] record
            for i, b in enumerate(baths):
                ct = time_means(b._cur_dev)[:Be]
                acc_c[0, i] += ct.sum()
                acc_c[1, i] += (ct * ct).sum()
            acc_e += time_means(mdrun._etot)[:Be].sum()
            tev[s][3].record(main)
    main.synchronize()
    side.synchronize()
    copy.synchronize()
    t1 = time.perf_counter()
    out = dict(power_sum=acc_p.cpu().numpy(), power2_sum=acc_p2.cpu().numpy(), cur_sum=acc_c[0].cpu().numpy(),
               cur_sq=acc_c[1].cpu().numpy(), ekin_sum=float(acc_e.item()))
    t2 = time.perf_counter()
    out["seconds"] = dict(total=t2 - t0, d2h=t2 - t1, setup=t0 - t_enter, overlap=overlap, sub_ensembles=nsub,
                          steps=sum(e[0].elapsed_time(e[1]) for e in tev) * 1e-3,
                          spectra=sum(e[1].elapsed_time(e[2]) for e in tev) * 1e-3,
                          reductions=sum(e[2].elapsed_time(e[3]) for e in tev) * 1e-3,
                          noise_gen=sum(e[0].elapsed_time(e[1]) for e in nev) * 1e-3,
                          note="noise_gen includes the upload of the spectral factors and, with overlap, runs beside steps")
    dw = 2. * N.pi / mdrun.dt / nmd
    mdrun.power_sum, mdrun.power2_sum = out["power_sum"], out["power2_sum"]
    mdrun.power = N.stack((dw * N.arange(nmd), out["power_sum"] / batch_total), axis=1)
    mdrun.power2 = N.stack((dw * N.arange(nmd), out["power2_sum"] / batch_total), axis=1)
    out["state"] = state
    return out
