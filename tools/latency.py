#!/usr/bin/env python
"""Small-batch latency of the reference's per-GP call pattern (GP(...), update_mll_and_dmll!, predict_y with one test
column) through the host API: B in {1, 4, 12} GPs of one trial at n = 2000.  Run on a B200."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpr_jl_b200 as G  # noqa: E402
from gpr_jl_b200 import data  # noqa: E402

rows = []
for system, B in (("CP", 1), ("CP", 4), ("FB", 12)):
    tr = data.make_config(system, trials=1, n_test=4)[0]
    gps = [G.GPE(tr["X"], tr["Y"][k], G.MeanZero(), G.SEArd(tr["theta0"][k][1:-1], tr["theta0"][k][-1]), logNoise=tr["theta0"][k][0])
           for k in range(B)]
    batch = G.GPBatch(gps)
    batch.eval(grad=True)

    def timeit(fn, reps=5):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3

    t_v = timeit(lambda: batch.eval(grad=False))
    x1 = tr["Xtest"][:, :1]
    t_p1 = timeit(lambda: batch.predict_y(x1, var=True))     # value-only state (after optimize!): blocked forward substitution
    t_pm = timeit(lambda: batch.predict_y(x1, var=False))
    t_g = timeit(lambda: batch.eval(grad=True))
    t_p1v = timeit(lambda: batch.predict_y(x1, var=True))    # V = L^-T resident (after a gradient evaluation): split GEMV path
    rows.append({"system": system, "B": B, "n": 2000, "d": tr["X"].shape[0], "eval_grad_ms": t_g, "eval_value_ms": t_v,
                 "predict_1col_meanvar_ms": t_p1, "predict_1col_meanvar_Vresident_ms": t_p1v, "predict_1col_mean_ms": t_pm})
    print(json.dumps(rows[-1]), flush=True)
    batch.close()
    del batch
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
