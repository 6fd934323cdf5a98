"""Generates tests/golden/golden_small.npz: seeded small inputs and the ORACLE's outputs for them.

Provenance: the reference pins no numbers on this path (no tests / fixtures / seeds; Julia absent), so these vectors are
restatement-derived - produced by oracle/gp_oracle.py (scipy LAPACK, reference operation order) - and labelled as
such.  They guard the oracle against regressions (CPU test) and give the CUDA path a fixed target (GPU test)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gpr_jl_b200  # noqa: E402,F401
from gpr_jl_b200 import data  # noqa: E402
from oracle import gp_oracle as go  # noqa: E402

CASES = [  # name, system, n, seed, kind, theta source
    ("p1_n40_se", "P1", 40, 11, "se", "rule"),
    ("p1_n130_se_cfg", "P1", 130, 12, "se", "P1_MAX64"),
    ("p2_n150_se", "P2", 150, 13, "se", "rule"),
    ("cp_n96_mat52", "CP", 96, 14, "mat52", "rule"),
    ("fb_n64_mat32", "FB", 64, 15, "mat32", "rule"),
    ("p1_n33_mat12", "P1", 33, 16, "mat12", "rule"),
]


def build():
    out = {}
    for name, system, n, seed, kind, src in CASES:
        tr = data.make_trial(system, n, seed, n_test=7)
        X = np.ascontiguousarray(tr["X"].T)
        th = data.theta0(system, tr["X"], None if src == "rule" else src)
        if src == "rule":
            th[1:-1] -= 1.0
        G = tr["Y"].shape[0]
        mll = np.zeros(G); grad = np.zeros((G, th.size)); mu = np.zeros((G, 7)); var = np.zeros((G, 7)); info = np.zeros(G, int)
        for k in range(G):
            r = go.eval_mll(X, tr["Y"][k], th, kind=kind, return_state=True)
            mll[k], grad[k], info[k] = r["mll"], r["grad"], r["info"]
            mu[k], var[k] = go.predict(X, th, r["state"], np.ascontiguousarray(tr["Xtest"].T), kind=kind)
        for key, val in dict(X=tr["X"], Y=tr["Y"], Xtest=tr["Xtest"], theta=th, mll=mll, grad=grad, mu=mu, var=var, info=info).items():
            out[f"{name}/{key}"] = val
        out[f"{name}/kind"] = np.array(kind)
    return out


if __name__ == "__main__":
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz")
    np.savez_compressed(dst, **build())
    print("wrote", dst, os.path.getsize(dst), "bytes")
