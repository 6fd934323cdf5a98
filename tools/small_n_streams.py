#!/usr/bin/env python
"""Stream-group count at small and medium n (run on a B200 with GPRB200_STREAMS=1|2|4|8): logML+gradient and value-only
evaluations/s at n = 256, 512, 1024 (d = 26, B = 512 / 380), device-resident timing.  Result: 2 groups
are 0-3 % ahead of the default 4, 1 and 8 are behind."""
import os, sys, json, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import gpr_jl_b200 as G
dev = torch.device("cuda", 0)
for n in (256, 512, 1024):
    d = 26; B = 512 if n < 1024 else 380
    rng = np.random.default_rng(n)
    th = np.concatenate([[-2.0], np.full(d, np.log(10.0)), [0.0]])
    gps = []
    for t in range(B // 4):
        X = np.asfortranarray(rng.standard_normal((d, n)))
        for k in range(4):
            y = np.sin(X[k]) + 0.1 * rng.standard_normal(n)
            gps.append(G.GPE(X, y, G.MeanZero(), G.SEArd(th[1:-1], th[-1]), logNoise=th[0]))
    batch = G.GPBatch(gps)
    P = d + 2
    thetas = [torch.from_numpy(np.tile(th, (B, 1)) + 0.05 * rng.standard_normal((B, P))).to(dev) for _ in range(4)]
    mll = torch.empty(B, dtype=torch.float64, device=dev); grad = torch.empty(B, P, dtype=torch.float64, device=dev)
    info = torch.empty(B, dtype=torch.int32, device=dev); st = torch.cuda.current_stream()
    for wg in (True, False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        batch.eval_device(thetas[0].data_ptr(), mll.data_ptr(), grad.data_ptr() if wg else None, info.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize(); e0.record(st)
        for r in range(8):
            batch.eval_device(thetas[(r + 1) % 4].data_ptr(), mll.data_ptr(), grad.data_ptr() if wg else None, info.data_ptr(), st.cuda_stream)
        e1.record(st); torch.cuda.synchronize()
        print(os.environ.get("GPRB200_STREAMS", "4"), n, "grad" if wg else "value", round(B / (e0.elapsed_time(e1) * 1e-3 / 8), 1), flush=True)
    batch.close()
