#!/usr/bin/env python
"""Synthetic sweep of BASELINE.json configs[4]: n in {256..8192} x d in {13,26,39,52}, one B200.

Per point: logML+gradient evaluations/s, value-only evaluations/s (the line-search trials of the optimiser) and
predict_y samples/s (mean+variance, m = 100 test columns per GP), with the fraction of the measured FP64 peak computed
from the ALGORITHMIC flop counts of SURVEY.md section 8d.  Inputs: iid N(0,1) (any d) with theta from the rule of
examples/maximal_coordinates/FBparam.jl:23-26 (s_f = 1, l_d = 10/std_d, logNoise = -2).  Device-resident timing (CUDA events).
Usage (on a B200):  python tools/sweep.py --out profiles/r01_sweep.json [--quick] [--cpu]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def f_eval(n, d):
    return n ** 3 + 4 * d * n ** 2 + 4 * n ** 2


def f_value(n, d):
    return n ** 3 / 3 + 1.5 * d * n ** 2 + 2 * n ** 2


def f_pred(n, d):
    return n ** 2 + (3 * d + 4) * n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true", help="n <= 2048 only")
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle (one evaluation per point, n <= 2048)")
    ap.add_argument("--gb", type=float, default=8.0, help="target HBM footprint per point")
    ap.add_argument("--nmin", type=int, default=0, help="skip sizes below this n")
    ap.add_argument("--nmax", type=int, default=1 << 30, help="skip sizes above this n")
    ap.add_argument("--dims", default="13,26,39,52", help="comma-separated input dimensions")
    ap.add_argument("--bmax", type=int, default=4096, help="cap on GPs per point (B is chosen to fill --gb of HBM)")
    a = ap.parse_args()
    import torch
    import gpr_jl_b200 as G
    peak = json.load(open(os.path.join(ROOT, "profiles", "FP64_PEAKS.json")))["dgemm_tflops_sustained"]
    dev = torch.device("cuda", 0)
    ns = [n for n in [256, 512, 1024, 2048] + ([] if a.quick else [4096, 8192]) if a.nmin <= n <= a.nmax]
    rows = []
    for n in ns:
        for d in [int(x) for x in a.dims.split(",")]:
            npad = (n + 127) // 128 * 128
            per_gp = 2 * npad * npad * 8 + 3 * npad * 128 * 8 + npad * 128 * 8
            B = int(max(4, min(a.bmax, a.gb * 1e9 // per_gp)))
            G4 = 4  # GPs per dataset (like the 4 outputs of a CP trial)
            B = B // G4 * G4
            rng = np.random.default_rng(n * 100 + d)
            gps = []
            th = np.concatenate([[-2.0], np.full(d, np.log(10.0)), [0.0]])
            for t in range(B // G4):
                X = np.asfortranarray(rng.standard_normal((d, n)))
                for k in range(G4):
                    y = np.sin(X[k % d]) + 0.1 * rng.standard_normal(n)
                    gps.append(G.GPE(X, y, G.MeanZero(), G.SEArd(th[1:-1], th[-1]), logNoise=th[0]))
            batch = G.GPBatch(gps)
            P = d + 2
            thetas = [torch.from_numpy(np.tile(th, (B, 1)) + 0.05 * rng.standard_normal((B, P))).to(dev) for _ in range(4)]
            mll = torch.empty(B, dtype=torch.float64, device=dev)
            grad = torch.empty(B, P, dtype=torch.float64, device=dev)
            info = torch.empty(B, dtype=torch.int32, device=dev)
            st = torch.cuda.current_stream()

            def run(with_grad, reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                batch.eval_device(thetas[0].data_ptr(), mll.data_ptr(), grad.data_ptr() if with_grad else None, info.data_ptr(), st.cuda_stream)
                torch.cuda.synchronize()
                e0.record(st)
                for r in range(reps):
                    batch.eval_device(thetas[(r + 1) % 4].data_ptr(), mll.data_ptr(), grad.data_ptr() if with_grad else None, info.data_ptr(), st.cuda_stream)
                e1.record(st)
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) * 1e-3 / reps

            reps = 3 if n >= 2048 else 6
            tg = run(True, reps)
            ok = int((info >= 0).sum().item())
            tv = run(False, reps)
            m = 100
            Xs = np.asfortranarray(rng.standard_normal((d, m)))
            batch.predict_y(Xs, var=True)
            tp = 0.0
            for _ in range(3):  # device time of gprb_predict (CUDA events on the library's stream, copies included)
                batch.predict_y(Xs, var=True)
                tp += batch.last_predict_ms(0) * 1e-3 / 3
            row = {"n": n, "d": d, "B": B, "info_ok": ok,
                   "evals_per_s": B / tg, "frac_fp64_peak": f_eval(n, d) * B / tg / 1e12 / peak,
                   "value_only_evals_per_s": B / tv, "value_only_frac_fp64_peak": f_value(n, d) * B / tv / 1e12 / peak,
                   "predict_samples_per_s": B * m / tp, "predict_frac_fp64_peak": f_pred(n, d) * B * m / tp / 1e12 / peak}
            if a.cpu and n <= 2048:
                from oracle import gp_oracle as go
                Xc = np.ascontiguousarray(gps[0].x.T)
                t0 = time.perf_counter()
                go.eval_mll(Xc, gps[0].y, th, with_grad=True)
                row["cpu_oracle_evals_per_s"] = 1.0 / (time.perf_counter() - t0)
                row["cpu_threads"] = os.cpu_count()
            rows.append(row)
            print(json.dumps(row), flush=True)
            batch.close()
            del batch, gps
    if a.out:
        json.dump({"fp64_peak_tflops": peak, "peak_source": "profiles/FP64_PEAKS.json cuBLAS DGEMM 8192^3 sustained",
                   "inputs": "iid N(0,1), theta = [-2, log 10 x d, 0] + 0.05 N(0,1)", "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
