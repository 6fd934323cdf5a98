#!/usr/bin/env python
"""predict_y throughput through gprb_predict (host buffers in and out) at the bench configuration: B GPs of the CP
system at n = 2000, m test columns per GP shared by the batch.  One sample = (mu*, var*) of one GP at one test column;
F_pred = n^2 + (3d + 4) n flops per sample.  Run on a B200: python tools/predict_bench.py [--trials T]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpr_jl_b200 as G  # noqa: E402
from gpr_jl_b200 import data  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--trials", type=int, default=100)
ap.add_argument("--n-train", dest="n", type=int, default=2000)
ap.add_argument("--system", default="CP")
ap.add_argument("--ms", default="100,128,64,200,1000")
ap.add_argument("--out", default=None)
args = ap.parse_args()

trials = data.make_config(args.system, trials=args.trials, n=args.n)
gps = []
for tr in trials:
    for k in range(tr["Y"].shape[0]):
        th = tr["theta0"][k]
        gps.append(G.GPE(tr["X"], tr["Y"][k], G.MeanZero(), G.SEArd(th[1:-1], th[-1]), logNoise=th[0]))
batch = G.GPBatch(gps)
batch.eval(grad=False)
B, n, d = batch.B, args.n, trials[0]["X"].shape[0]
try:
    peak = json.load(open(os.path.join(ROOT, "profiles", "FP64_PEAKS.json")))["dgemm_tflops_sustained"]
except Exception:
    peak = 35.43
rows = []
for m in [int(x) for x in args.ms.split(",")]:
    Xt = data.make_trial(args.system, 8, seed=99, n_test=m)["Xtest"]
    batch.predict_y(Xt, var=True)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        batch.predict_y(Xt, var=True)
    tv = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        batch.predict_y(Xt, var=False)
    tm = (time.perf_counter() - t0) / reps
    f = n * n + (3 * d + 4) * n
    row = {"B": B, "n": n, "d": d, "m": m, "mean_var_ms": tv * 1e3, "mean_only_ms": tm * 1e3,
           "samples_per_s_mean_var": B * m / tv, "samples_per_s_mean_only": B * m / tm,
           "frac_fp64_peak_mean_var": B * m / tv * f / 1e12 / peak}
    rows.append(row)
    print(json.dumps(row), flush=True)
if args.out:
    json.dump(rows, open(args.out, "w"), indent=1)
