#!/usr/bin/env python
"""Stage breakdown of one logML+gradient evaluation at small n (profiled pass: one stream, events around every stage and
every tile-GEMM launch) next to the unprofiled multi-stream time.  Run on a B200: python tools/small_n_breakdown.py"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import gpr_jl_b200 as G
dev = torch.device("cuda", 0)
out = []
for n, B in ((256, 3812), (512, 1268), (1024, 380)):
    d = 26
    rng = np.random.default_rng(n)
    th = np.concatenate([[-2.0], np.full(d, np.log(10.0)), [0.0]])
    gps = []
    for t in range(B // 4):
        X = np.asfortranarray(rng.standard_normal((d, n)))
        for k in range(4):
            y = np.sin(X[k]) + 0.1 * rng.standard_normal(n)
            gps.append(G.GPE(X, y, G.MeanZero(), G.SEArd(th[1:-1], th[-1]), logNoise=th[0]))
    B = len(gps)
    batch = G.GPBatch(gps)
    P = d + 2
    thetas = [torch.from_numpy(np.tile(th, (B, 1)) + 0.05 * rng.standard_normal((B, P))).to(dev) for _ in range(4)]
    mll = torch.empty(B, dtype=torch.float64, device=dev); grad = torch.empty(B, P, dtype=torch.float64, device=dev)
    info = torch.empty(B, dtype=torch.int32, device=dev); st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    batch.eval_device(thetas[0].data_ptr(), mll.data_ptr(), grad.data_ptr(), info.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize(); e0.record(st)
    for r in range(8):
        batch.eval_device(thetas[(r + 1) % 4].data_ptr(), mll.data_ptr(), grad.data_ptr(), info.data_ptr(), st.cuda_stream)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    batch.set_profiling(True)
    for r in range(2):
        batch.eval_device(thetas[r].data_ptr(), mll.data_ptr(), grad.data_ptr(), info.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    stage = batch.last_stage_ms(); launches = batch.last_gemm_launch_ms()
    batch.set_profiling(False)
    rec = {"n": n, "B": B, "ms_per_eval_batch": ms, "stage_ms_profiled": {k: round(v, 3) for k, v in stage.items()},
           "gemm_launch_ms": [(m, s, round(t, 3)) for m, s, t in launches]}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    batch.close()
