#!/usr/bin/env python
"""Per-launch breakdown of the tile GEMM at the bench configuration: device time of every k_tile_gemm launch of one
profiled evaluation, the DMMA flops that launch executes (a host replica of the kernel's chunk/skip logic), and the
resulting executed-TFLOP/s against the measured DMMA peak.  Run on a B200: python tools/gemm_breakdown.py [--trials T]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NB, KT = 128, 16


def warp_dmma(mode, c, wm, wn, ragged, mi_valid, first_kb_chunks, sym=False):
    """8x8x4 DMMAs one consumer warp issues for main-loop chunk c (4 k4-steps x selected (mi, ni) blocks)."""
    mi_lim, ni_lim, off = mi_valid, 4, -64
    if not ragged:
        if mode == "chol_diag":
            off = 4 * wn - 8 * wm
        elif mode == "lauum" and c < first_kb_chunks:
            mi_lim = max(0, min(8, 2 * c + 2 - 8 * wm))
        elif mode == "lauum" and sym:
            off = 4 * wn - 8 * wm
        elif mode == "trtri_row" and c < NB // KT:
            ni_lim = max(0, min(4, 2 * c + 2 - 4 * wn))
    cnt = sum(1 for mi in range(mi_lim) for ni in range(ni_lim) if mi >= ni + off)
    return 4 * cnt


def launch_flops(mode, step, n):
    J = (n + NB - 1) // NB
    nv = (n + KT - 1) // KT * KT
    nvl = nv - (J - 1) * NB
    last_chunks = nvl // KT
    if mode == "chol_diag":
        tiles = [(step, step, 0, step, 0)]
    elif mode == "chol_col":
        tiles = [(i, step, 0, step, 1) for i in range(step + 1, J)]
    elif mode == "trtri_row":
        tiles = [(step, j, j, step, 2) for j in range(step)]
    else:
        tiles = [(i, j, i, J, 0) for i in range(J) for j in range(i + 1)]
    total = 0
    for (i, j, kb0, kb1, post) in tiles:
        nchunks = (kb1 - kb0) * (NB // KT) - ((NB // KT - last_chunks) if (kb1 == J and kb1 > kb0) else 0)
        rows_valid = nvl if i == J - 1 else NB
        ragged = rows_valid != NB
        first_kb_chunks = min(nchunks, NB // KT)
        for w in range(8):
            s4, h = w & 3, w >> 2
            wm, wn = h, (3 - s4) if h else s4
            mi_valid = min(8, max(0, (rows_valid - wm * 64 + 7) // 8)) if ragged else 8
            for c in range(nchunks):
                total += warp_dmma(mode, c, wm, wn, ragged, mi_valid, first_kb_chunks, sym=(i == j))
            if post:
                kmax = (wn * 32 + 31) if post == 1 else min(wm * 64 + 63, rows_valid - 1)
                for c in range(NB // KT):
                    if c * KT <= kmax:
                        mi_n, ni_n = mi_valid, 4
                        if not ragged:  # triangular skip inside the warp tile (SEL_NLO2 / SEL_MLOx)
                            if post == 1 and 2 * c - 4 * wn == 2:
                                ni_n = 2
                            if post == 2 and 2 * c - 8 * wm > 0:
                                mi_n = 8 - (2 * c - 8 * wm)
                        total += 4 * mi_n * ni_n
    return total * 512  # 8x8x4 MACs x 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=100)
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import gpr_jl_b200 as G
    from gpr_jl_b200 import data
    trials = data.make_config("CP", trials=a.trials, n=a.n)
    gps = [G.GPE(tr["X"], tr["Y"][k], G.MeanZero(), G.SEArd(tr["theta0"][k][1:-1], tr["theta0"][k][-1]), logNoise=tr["theta0"][k][0])
           for tr in trials for k in range(tr["Y"].shape[0])]
    batch = G.GPBatch(gps)
    B = batch.B
    batch.eval(grad=True)
    batch.set_profiling(True)
    batch.eval(grad=True)
    batch.eval(grad=True)
    st = batch.last_stage_ms()
    rows = batch.last_gemm_launch_ms()
    peak = json.load(open(os.path.join(ROOT, "profiles", "FP64_PEAKS.json")))["micro"]["dmma_tflops_1cta"]
    agg = {}
    out = []
    for mode, step, ms in rows:
        fl = launch_flops(mode, step, a.n) * B
        tf = fl / (ms * 1e-3) / 1e12
        out.append({"mode": mode, "step": step, "ms": ms, "executed_gflop": fl / 1e9, "executed_tflops": tf})
        m = agg.setdefault(mode, [0.0, 0.0])
        m[0] += ms
        m[1] += fl
    print(f"B={B} n={a.n}  stage_ms={ {k: round(v, 2) for k, v in st.items()} }")
    for r in out:
        print(f"{r['mode']:10s} step {r['step']:2d}  {r['ms']:8.3f} ms  {r['executed_gflop']:9.1f} GF executed  {r['executed_tflops']:6.2f} TF/s  ({r['executed_tflops'] / peak:5.1%} of DMMA peak {peak})")
    print("---- per mode")
    tot_ms = tot_fl = 0.0
    for mode, (ms, fl) in agg.items():
        print(f"{mode:10s} {ms:8.2f} ms  {fl / 1e12:7.3f} TF executed  {fl / ms / 1e9:6.2f} TF/s")
        tot_ms += ms
        tot_fl += fl
    alg = B * float(a.n) ** 3
    print(f"all GEMM   {tot_ms:8.2f} ms  executed {tot_fl / 1e12:.3f} TF ({tot_fl / tot_ms / 1e9:.2f} TF/s), algorithmic n^3 {alg / 1e12:.3f} TF ({alg / tot_ms / 1e9:.2f} TF/s), executed/algorithmic {tot_fl / alg:.3f}")
    if a.out:
        json.dump({"B": B, "n": a.n, "stage_ms": st, "launches": out, "dmma_peak_tflops": peak}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
