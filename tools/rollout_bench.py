#!/usr/bin/env python
"""BASELINE config 4 - "FB fourbar maximal-coordinate GP with implicitProjection on predictions, n=2000, 100 trials":
the batched, overlapped rollout of examples/utils/predictdynamics.jl:7-22 for all trials of one GPU.

Per trial: 12 GPs (FBnoise.jl:24), 100 test states, 20 steps (examples/noise.jl:62).  Every step predicts the 12 next-step
velocity components of all states of all trials (mean only - the reference discards the variance, predictdynamics.jl:13)
and hands them to the host, where `projectv!` (src/projections/implicitProjection.jl:80-107: Newton iterations on a dense
KKT system of size 6N + constraints <= ~46 per state) and `updatestate!` produce the next states.  ConstrainedDynamics.jl
is not available here, so the host step is a stand-in of the same shape and cost class: per state a few Newton iterations,
each one dense 40 x 40 solve (batched over the 100 states of a trial with numpy).

Reported: device-only time (host step = identity), host-only time, serial loop (predict -> wait -> host, one group),
overlapped loop (two trial groups alternating on the two prediction pipelines of gprb_predict_async).
Usage (on a B200):  python tools/rollout_bench.py [--system FB] [--trials 100] [--n-train 2000] [--out profiles/r02_rollout.json]"""
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
    ap.add_argument("--system", default="FB")
    ap.add_argument("--trials", type=int, default=100)
    ap.add_argument("--n-train", dest="n", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--states", type=int, default=100)
    ap.add_argument("--newton", type=int, default=3, help="Newton iterations of the projectv! stand-in")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import gpr_jl_b200 as G
    from gpr_jl_b200 import data, experiment

    trials = data.make_config(a.system, trials=a.trials, n=a.n, n_test=a.states)
    Gout = trials[0]["Y"].shape[0]
    d = trials[0]["X"].shape[0]
    gps = []
    for tr in trials:
        for k in range(Gout):
            th = tr["theta0"][k]
            gps.append(G.GPE(tr["X"], tr["Y"][k], G.MeanZero(), G.SEArd(th[1:-1], th[-1]), logNoise=th[0]))
    batch = G.GPBatch(gps)
    mll, _, info = batch.eval(grad=False)  # the state optimize! leaves behind: factor + alpha, no inverse
    idx = np.asarray(data.SYSTEMS[a.system]["outputs"]) - 1
    rng = np.random.default_rng(0)
    nk = 40  # 6 N + constraint rows of the fourbar KKT system
    Fk = rng.standard_normal((nk, nk)) + nk * np.eye(nk)

    def project(t, states, mu):  # stand-in for getvomega + projectv! + updatestate!
        nxt = states.copy()
        nxt[idx, :] = mu
        m = states.shape[1]
        v = np.zeros((m, nk))
        v[:, :idx.size] = mu.T
        for _ in range(a.newton):  # Newton: F \\ f per state, F re-assembled per iteration (a rank-1 change stands in for it)
            F = Fk[None, :, :] + 1e-3 * v[:, :, None] * v[:, None, :]
            v = v - 1e-3 * np.linalg.solve(F, v[:, :, None])[:, :, 0]
        nxt[idx, :] = v[:, :idx.size].T
        nxt[1, :] += 0.01 * nxt[8, :]
        return nxt

    def identity(t, states, mu):
        nxt = states.copy()
        nxt[idx, :] = mu
        return nxt

    starts = [tr["Xtest"] for tr in trials]
    res = {"system": a.system, "trials": a.trials, "n": a.n, "d": d, "G": Gout, "B": batch.B, "steps": a.steps, "states": a.states,
           "info_ok": int((info >= 0).sum())}
    experiment.predictdynamics(batch, Gout, starts, 2, identity, overlap=True)  # warm-up (scratch allocation)
    tm = {}
    experiment.predictdynamics(batch, Gout, starts, a.steps, identity, overlap=False, timing=tm)
    res["device_only_s"] = tm["total_s"]
    res["device_only_wait_s"] = tm["wait_s"]
    t0 = time.perf_counter()
    for _ in range(a.steps):
        for t in range(a.trials):
            project(t, starts[t], np.zeros((Gout, a.states)))
    res["host_only_s"] = time.perf_counter() - t0
    tm = {}
    fs = experiment.predictdynamics(batch, Gout, starts, a.steps, project, overlap=False, timing=tm)
    res["serial_s"], res["serial_wait_s"], res["serial_host_s"] = tm["total_s"], tm["wait_s"], tm["host_s"]
    tm = {}
    fo = experiment.predictdynamics(batch, Gout, starts, a.steps, project, overlap=True, timing=tm)
    res["overlapped_s"], res["overlapped_wait_s"], res["overlapped_host_s"] = tm["total_s"], tm["wait_s"], tm["host_s"]
    res["identical_results"] = bool(all(np.array_equal(x, y) for x, y in zip(fs, fo)))
    res["bound_s"] = max(res["device_only_s"], res["host_only_s"])
    res["overlapped_over_bound"] = res["overlapped_s"] / res["bound_s"]
    res["predictions_per_s_overlapped"] = batch.B * a.states * a.steps / res["overlapped_s"]
    res["predictions_per_s_device_only"] = batch.B * a.states * a.steps / res["device_only_s"]
    print(json.dumps(res), flush=True)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
