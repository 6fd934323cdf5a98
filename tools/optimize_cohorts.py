#!/usr/bin/env python
"""Batched optimiser with the GPs split into K independent cohorts (one GPBatch each), optimised CONCURRENTLY from K host
threads (the C ABI allows different batches on different threads; ctypes drops the GIL during the call).  A cohort
advances in lock-step rounds; while one cohort is in a thin, latency-bound round (a few GPs deep in a line search or a
jitter retry) the others keep the GPU busy.  Usage (B200): python tools/optimize_cohorts.py [--cohorts 1,2,4] [--out f.json]"""
import argparse
import json
import os
import sys
import threading
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
ap.add_argument("--cohorts", default="1,2,4")
ap.add_argument("--out", default=None)
a = ap.parse_args()
trials = data.make_config("CP", trials=a.trials, n=a.n)
rows = []
ref = None
for K in [int(x) for x in a.cohorts.split(",")]:
    batches = []
    for c in range(K):
        gps = []
        for tr in trials[c::K]:
            th = data.theta0("CP", tr["X"])
            for k in range(tr["Y"].shape[0]):
                gps.append(G.GPE(tr["X"], tr["Y"][k], G.MeanZero(), G.SEArd(th[1:-1], th[-1]), logNoise=th[0]))
        b = G.GPBatch(gps)
        b.eval(grad=False)
        batches.append(b)
    results = [None] * K

    def work(c):
        results[c] = batches[c].optimize(G.LBFGS(linesearch=G.BackTracking(order=2)), G.Options(iterations=a.iterations))

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(c,)) for c in range(K)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    f = sum(r["f_calls"] for res in results for r in res)
    fg = sum(r["g_calls"] for res in results for r in res)
    mins = {}
    for c in range(K):
        for i, r in enumerate(results[c]):
            mins[(c + K * (i // 4), i % 4)] = r["minimum"]
    if ref is None:
        ref = mins
    same = all(mins[k] == ref[k] for k in ref)
    row = {"cohorts": K, "B": sum(b.B for b in batches), "seconds": dt, "value_only_evals": f, "value_grad_evals": fg,
           "evals_per_s": (f + fg) / dt, "minima_identical_to_one_cohort": bool(same)}
    rows.append(row)
    print(json.dumps(row), flush=True)
    for b in batches:
        b.close()
if a.out:
    json.dump(rows, open(a.out, "w"), indent=1)
