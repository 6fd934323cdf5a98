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
