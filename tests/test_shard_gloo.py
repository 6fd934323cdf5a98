"""World-size-2 gloo test of the trial sharding + gather host logic (no GPU): each rank evaluates its trials with the
CPU oracle standing in for the device evaluation and the gathered table must equal the single-process table."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import gpr_jl_b200  # noqa: F401
from gpr_jl_b200 import data, shard
from oracle import gp_oracle as go

N_TRIALS, N = 5, 40


def _table_row(t):
    tr = data.make_trial("P1", N, seed=500 + t)
    th = data.theta0("P1", tr["X"])
    X = np.ascontiguousarray(tr["X"].T)
    return np.array([go.eval_mll(X, tr["Y"][k], th, with_grad=False)["mll"] for k in range(3)])


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.trials_for_rank(N_TRIALS, rank, world)
    local = {t: _table_row(t) for t in mine}
    out = shard.gather_trial_results(local, N_TRIALS, 3)
    q.put((rank, mine, out))
    dist.destroy_process_group()


def test_round_robin_partition():
    assert shard.trials_for_rank(10, 1, 4) == [1, 5, 9]
    owned = sorted(t for r in range(8) for t in shard.trials_for_rank(100, r, 8))
    assert owned == list(range(100))
    assert all(shard.owner_of(t, 8) == t % 8 for t in range(100))


def test_gather_world2_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.stack([_table_row(t) for t in range(N_TRIALS)])
    for rank, mine, out in got:
        assert mine == list(range(rank, N_TRIALS, 2))
        np.testing.assert_array_equal(out, ref)


def test_gather_without_process_group_is_identity():
    out = shard.gather_trial_results({0: np.array([1.0, 2.0]), 1: np.array([3.0, 4.0])}, 2, 2)
    np.testing.assert_array_equal(out, [[1, 2], [3, 4]])


def test_gp_level_partition_is_balanced_and_complete():
    """shard.gps_for_rank: contiguous (trial, output) ranges whose sizes differ by at most one and that cover every GP once."""
    for T, G, W in ((100, 4, 8), (100, 12, 8), (5, 3, 2), (7, 4, 3), (3, 3, 4)):
        sizes, seen = [], []
        for r in range(W):
            units = shard.gps_for_rank(T, G, r, W)
            assert all(0 <= g0 < g1 <= G for _, g0, g1 in units)
            sizes.append(sum(g1 - g0 for _, g0, g1 in units))
            seen += [(t, g) for t, g0, g1 in units for g in range(g0, g1)]
        assert seen == [(t, g) for t in range(T) for g in range(G)]
        assert max(sizes) - min(sizes) <= 1
    assert [sum(g1 - g0 for _, g0, g1 in shard.gps_for_rank(100, 4, r, 8)) for r in range(8)] == [50] * 8


def _worker_gp(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units = shard.gps_for_rank(N_TRIALS, 3, rank, world)
    local = {}
    for (t, g0, g1) in units:  # a trial that straddles the two ranks is evaluated (its X "uploaded") on both, outputs disjoint
        row = _table_row(t)
        for g in range(g0, g1):
            local[t * 3 + g] = row[g:g + 1]
    out = shard.gather_trial_results(local, N_TRIALS * 3, 1)  # rows keyed by GP
    q.put((rank, units, out))
    dist.destroy_process_group()


def test_gather_world2_gp_level_split_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_gp, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.stack([_table_row(t) for t in range(N_TRIALS)]).reshape(-1, 1)
    straddling = [t for t in range(N_TRIALS) if sum(1 for _, units, _ in got for (u, _, _) in units if u == t) == 2]
    assert straddling == [2]  # 15 GPs over 2 ranks: 7 + 8, trial 2 is cut after its first output
    for rank, units, out in got:
        np.testing.assert_array_equal(out, ref)
