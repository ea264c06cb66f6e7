"""Multi-rank path on CPU (gloo, world_size 2): ensemble sharding and the single all-reduce that
forms the ensemble averages.  No GPU: the reduction plumbing is exercised with
synthetic per-rank partial sums, and the result must equal the single-process average."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, btotal, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pymd_b200 import ensemble
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        t0, n = ensemble.shard(btotal)
        # per-trajectory "results" that depend only on the GLOBAL trajectory id, as the Philox streams do
        ids = np.arange(t0, t0 + n)
        power = np.stack([np.sin(0.1 * ids) ** 2 * (k + 1) for k in range(8)], 0).sum(1)      # sum over local trajectories
        cur = np.array([np.cos(0.3 * ids).sum(), (0.01 * ids).sum()])
        cur2 = np.array([(np.cos(0.3 * ids) ** 2).sum(), ((0.01 * ids) ** 2).sum()])
        res = ensemble.allreduce_sums([power, cur, cur2, np.array([float(n)])])
        out.put((rank, t0, n, [r.tolist() for r in res]))
    finally:
        dist.destroy_process_group()


def test_shard_is_a_contiguous_partition():
    from pymd_b200 import ensemble
    for total in (1, 7, 4096, 8191):
        for world in (1, 2, 3, 8):
            parts = [ensemble.shard(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(n for _, n in parts) == total
            for (a0, an), (b0, _) in zip(parts, parts[1:]):
                assert a0 + an == b0
            assert max(n for _, n in parts) - min(n for _, n in parts) <= 1
    assert ensemble.shard(10) == (0, 10)                       # not distributed -> whole ensemble


def test_allreduce_identity_when_not_distributed():
    from pymd_b200 import ensemble
    a = [np.arange(4.0), np.array([[1.0, 2.0]])]
    r = ensemble.allreduce_sums(a)
    assert all(np.array_equal(x, y) for x, y in zip(a, r)) and r[0] is not a[0]


@pytest.mark.timeout(120)
def test_two_rank_gloo_average_equals_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port, btotal = _free_port(), 37
    procs = [ctx.Process(target=_worker, args=(r, 2, port, btotal, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert [g[1:3] for g in got] == [(0, 19), (19, 18)]
    ids = np.arange(btotal)
    power = np.stack([np.sin(0.1 * ids) ** 2 * (k + 1) for k in range(8)], 0).sum(1)
    cur = np.array([np.cos(0.3 * ids).sum(), (0.01 * ids).sum()])
    for g in got:                                               # both ranks hold the same global sums
        assert np.allclose(g[3][0], power, rtol=1e-13) and np.allclose(g[3][1], cur, rtol=1e-13)
        assert g[3][3] == [float(btotal)]
    assert got[0][3] == got[1][3]                               # bitwise identical on every rank
