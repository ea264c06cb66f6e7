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
    cur_t = [b._cur_dev[:, :mdrun.batch].mean(0) for b in mdrun.baths]          # time mean per trajectory
    cur_sum = N.array([float(c.sum().item()) for c in cur_t])
    cur_sq = N.array([float((c * c).sum().item()) for c in cur_t])
    ek = float(mdrun._etot[:, :mdrun.batch].mean(0).sum().item())
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
