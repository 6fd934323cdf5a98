#!/usr/bin/env python
"""End-to-end throughput of the batched optimiser (gprb_optimize = the reference's optimize! for every GP of the CP
configuration at once): B = 400 GPs, n = 2000, a fixed number of L-BFGS iterations.  Reports evaluations/s inside the
optimiser (value-only line-search trials + value+gradient evaluations) and the improvement of the objective.
Usage (B200): python tools/optimize_bench.py [--trials 100] [--iterations 3] [--out file.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpr_jl_b200 as G  # noqa: E402
from gpr_jl_b200 import data  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--trials", type=int, default=100)
ap.add_argument("--n", type=int, default=2000)
ap.add_argument("--iterations", type=int, default=3)
ap.add_argument("--out", default=None)
a = ap.parse_args()
trials = data.make_config("CP", trials=a.trials, n=a.n)
gps = []
for tr in trials:
    th = data.theta0("CP", tr["X"])  # rule-based start (FBparam.jl:23-26), a point the optimiser can improve on
    for k in range(tr["Y"].shape[0]):
        gps.append(G.GPE(tr["X"], tr["Y"][k], G.MeanZero(), G.SEArd(th[1:-1], th[-1]), logNoise=th[0]))
batch = G.GPBatch(gps)
mll0, _, _ = batch.eval(grad=False)
t0 = time.perf_counter()
res = batch.optimize(G.LBFGS(linesearch=G.BackTracking(order=2)), G.Options(iterations=a.iterations))
dt = time.perf_counter() - t0
f = sum(r["f_calls"] for r in res)
fg = sum(r["g_calls"] for r in res)
mll1 = np.array([-r["minimum"] for r in res])
out = {"B": batch.B, "n": a.n, "iterations": a.iterations, "seconds": dt, "value_only_evals": f, "value_grad_evals": fg,
       "evals_per_s": (f + fg + batch.B) / dt,  # + the final update_target! evaluation at the minimiser
       "mll_start_mean": float(mll0.mean()), "mll_end_mean": float(mll1.mean()), "improved": int((mll1 > mll0).sum()),
       "info_ok": int(sum(r["info"] >= 0 for r in res)), "ls_failed": int(sum(r["ls_failed"] for r in res))}
print(json.dumps(out))
if a.out:
    json.dump(out, open(a.out, "w"), indent=1)
