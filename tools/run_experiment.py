#!/usr/bin/env python
"""The reference's noise experiment for one system, end to end and batched (examples/noise.jl:100 ->
parallel/core.jl:28 -> maximal_coordinates/CPnoise.jl:12-54), on 1..8 GPUs:

    trials sharded round-robin over the ranks (shard.trials_for_rank)       <- Threads.@threads for jobid
    per rank: ONE GPBatch of all its trials' GPs, optimize! in lock-step       <- GP(...) + optimize! per output
    20-step rollouts of every test state, one predict call per step            <- predictdynamics
    one all-gather of the per-trial rows (theta*, mll, info, k-step MSE)       <- lock-guarded result callbacks

The physics between two prediction steps (getvomega -> projectv! -> updatestate!, ConstrainedDynamics.jl) stays a
host callback; here it is the synthetic generator's own one-step map, so the rollout error is meaningful.
Run:  python tools/run_experiment.py --trials 16 --n-train 512                      (one GPU)
      torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_experiment.py --trials 16 --n-train 512
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="CP")
    ap.add_argument("--trials", type=int, default=16)
    ap.add_argument("--n-train", dest="n", type=int, default=512)
    ap.add_argument("--tests", type=int, default=100, help="test states per trial (examples/noise.jl:62 testsamples)")
    ap.add_argument("--steps", type=int, default=20, help="rollout steps (simsteps)")
    ap.add_argument("--iterations", type=int, default=5)
    ap.add_argument("--time-limit", type=float, default=0.0,
                    help="Optim.Options(time_limit=...) per GP on the deterministic virtual clock (CPnoise.jl:41 uses 10 s); 0 = off")
    ap.add_argument("--cost-value", type=float, default=1.3, help="virtual seconds charged per value-only evaluation")
    ap.add_argument("--cost-grad", type=float, default=3.3, help="virtual seconds charged per value+gradient evaluation "
                    "(defaults: one host core of the bench box at n = 2000, from the cpu_baseline of bench.py)")
    ap.add_argument("--theta-key", default=None, help="start point from the reference's config.json entry (e.g. CP_MAX2048) instead of the rule")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import gpr_jl_b200 as G
    from gpr_jl_b200 import data, experiment, shard

    mine = shard.trials_for_rank(a.trials, rank, world)
    spec = data.SYSTEMS[a.system]
    idx = np.asarray(spec["outputs"]) - 1
    trials = [data.make_trial(a.system, a.n, seed=9000 + t, n_test=a.tests) for t in mine]
    # one shared start point for every GP, like the reference's config.json entry: derived from trial 0 on every rank
    ref_X = data.make_trial(a.system, a.n, seed=9000)["X"]
    th0 = data.theta0(a.system, ref_X, a.theta_key)
    params = np.exp(np.concatenate([[th0[-1]], th0[1:-1]]))  # config.json order [s_f, l...]
    t0 = time.perf_counter()
    if a.time_limit > 0:  # the reference's stopping rule: a per-GP time budget, here on a virtual clock
        opts = G.Options(iterations=1000, time_limit=a.time_limit, cost_value=a.cost_value, cost_grad=a.cost_grad)
    else:
        opts = G.Options(iterations=a.iterations)
    batch, res = experiment.fit_trials(trials, params, options=opts)
    t_fit = time.perf_counter() - t0
    Gout = idx.size

    def step_fn(t, states, mu):  # stand-in for getvomega + projectv! + updatestate!: write the predicted velocities back
        nxt = states.copy()
        nxt[idx, :] = mu
        return nxt

    t0 = time.perf_counter()
    final = experiment.predictdynamics(batch, Gout, [tr["Xtest"] for tr in trials], a.steps, step_fn)
    t_roll = time.perf_counter() - t0
    P = trials[0]["X"].shape[0] + 2
    rows = {}
    for k, t in enumerate(mine):
        mse = float(np.mean((final[k][idx, :] - trials[k]["Ytest"]) ** 2))  # 1-step targets of the generator as a proxy
        th = np.concatenate([r["minimizer"] for r in res[k * Gout:(k + 1) * Gout]])
        mll = [-r["minimum"] for r in res[k * Gout:(k + 1) * Gout]]
        info = [float(r["info"]) for r in res[k * Gout:(k + 1) * Gout]]
        rows[t] = np.concatenate([th, mll, info, [mse]])
    table = shard.gather_trial_results(rows, a.trials, Gout * P + 2 * Gout + 1, device=torch.device("cuda", local))
    if rank == 0:
        print(json.dumps({"system": a.system, "trials": a.trials, "n": a.n, "world": world, "gps_per_rank": batch.B,
                          "fit_seconds": t_fit, "rollout_seconds": t_roll,
                          "optimizer": {"time_limit_per_gp": a.time_limit, "cost_value": a.cost_value, "cost_grad": a.cost_grad} if a.time_limit > 0 else {"iterations": a.iterations},
                          "mean_iterations": float(np.mean([r["iterations"] for r in res])), "evaluations": int(sum(r["f_calls"] + r["g_calls"] for r in res)), "rollout_predictions": len(mine) * Gout * a.tests * a.steps,
                          "gathered_rows": int(np.isfinite(table[:, -1]).sum()), "mean_kstep_mse": float(np.mean(table[:, -1])), "mean_mll": float(np.mean(table[:, Gout * P:Gout * P + Gout])),
                          "all_info_ok": bool(np.all(table[:, Gout * P + Gout:Gout * P + 2 * Gout] >= 0))}))
    batch.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
