"""Trial sharding and the final gather.

The reference's only data-parallel axis is the trial index: ``Threads.@threads for jobid in 1:nruns``
(/root/reference/examples/parallel/core.jl:28), every trial using its own dataset
(examples/maximal_coordinates/CPnoise.jl:13-17).  Here one process drives one GPU (torchrun), trial t lives on rank
``t mod world`` with all G GPs of the trial co-located (they share X), nothing is exchanged while optimising or
predicting, and one all-gather of the small per-trial results (theta*, mll, info, predictions) replaces the
lock-guarded result callbacks of core.jl:47-56.

On GPUs the gather is ``gprb_gather`` - one ``ncclAllGather`` inside libgprb200.so on the library's own communicator
(``gp.context().comm_init()``), the entry point a Julia host binds as well.  The torch.distributed form below is the
host-logic twin used by the world_size-2 gloo tests on CPU (same padding / scatter-by-trial-id scheme).
"""
from __future__ import annotations

import numpy as np


def trials_for_rank(n_trials: int, rank: int, world: int):
    """Round-robin: trial t -> rank t mod world."""
    return list(range(rank, n_trials, world))


def owner_of(trial: int, world: int) -> int:
    return trial % world


def gps_for_rank(n_trials: int, outputs_per_trial: int, rank: int, world: int):
    """Balanced split at GP granularity: the n_trials * G GPs, in trial-major order, are cut into ``world`` contiguous
    ranges whose sizes differ by at most one.  Returns [(trial, first output, last output + 1), ...] for this rank.

    The G GPs of a trial are independent problems that only share the read-only inputs X, so a trial may straddle two
    ranks (both upload its X).  With 100 trials on 8 GPUs the trial split gives the busiest rank 13 trials against a mean
    of 12.5 (4 % imbalance); the GP split gives every rank exactly 50 (CP) / 150 (FB) GPs."""
    total = n_trials * outputs_per_trial
    lo, hi = total * rank // world, total * (rank + 1) // world
    out = []
    for t in range(lo // outputs_per_trial, (hi + outputs_per_trial - 1) // outputs_per_trial):
        g0, g1 = max(lo, t * outputs_per_trial) - t * outputs_per_trial, min(hi, (t + 1) * outputs_per_trial) - t * outputs_per_trial
        if g1 > g0:
            out.append((t, g0, g1))
    return out


def gather_trial_results(local: dict, n_trials: int, width: int, device=None, ctx=None):
    """All-gather per-trial result rows.  ``ctx``: a gp.context() whose comm_init() ran -> gprb_gather (NCCL inside
    the library); otherwise torch.distributed (gloo on CPU).

    local: {trial index -> 1-D float64 array of length ``width``} for the trials this rank owns.
    Returns an (n_trials, width) array on every rank, rows in trial order.  Equal-count padding keeps it a single
    fixed-size collective (latency-bound: <= ~50 KB per trial)."""
    if ctx is not None and getattr(ctx, "_comm", False):
        return ctx.gather(local, n_trials, width)
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = np.full((n_trials, width), np.nan)
        for t, row in local.items():
            out[t] = row
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (n_trials + world - 1) // world
    buf = torch.full((per, width + 1), float("nan"), dtype=torch.float64)
    for k, t in enumerate(sorted(local)):
        buf[k, 0] = float(t)
        buf[k, 1:] = torch.from_numpy(np.asarray(local[t], dtype=np.float64))
    if device is not None:
        buf = buf.to(device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = np.full((n_trials, width), np.nan)
    for p in parts:
        p = p.cpu().numpy()
        for row in p:
            if np.isfinite(row[0]):
                out[int(row[0])] = row[1:]
    return out
