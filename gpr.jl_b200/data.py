"""Synthetic CState-shaped GP workloads and the theta_0 fixtures.

The reference's datasets are absent (``*.jls`` is git-ignored, /root/reference/.gitignore:4) and cannot be
regenerated without ConstrainedDynamics.jl, so the workloads are synthesised with the same *shape*:
maximal-coordinate states of N planar bodies, 13 numbers per body in the CState order
``[x(3), q=(w,x,y,z)(4), v(3), omega(3)]`` (/root/reference/src/CState.jl:20), rotating about the x axis so that
per body only x_y, x_z, q_w, q_x, v_y, v_z, omega_x vary (the other six rows are structurally constant - the
constant-dimension edge case the reference's data always contains).  Targets are next-step velocity components
at the reference's output index sets.  Everything is seeded.
"""
from __future__ import annotations

import json
import os

import numpy as np

DT = 0.01  # GP step (examples/parallel/dataframes.jl:13)

# system -> (bodies, 1-based output indices, cart body index or None)
SYSTEMS = {
    "P1": dict(bodies=1, outputs=[9, 10, 11], cart=None),                                   # P1noise.jl:26
    "P2": dict(bodies=2, outputs=[9, 10, 22, 23, 11, 24], cart=None),                       # P2noise.jl:25
    "CP": dict(bodies=2, outputs=[9, 22, 23, 24], cart=0),                                  # CPnoise.jl:28
    "FB": dict(bodies=4, outputs=[9, 10, 22, 23, 35, 36, 48, 49, 11, 24, 37, 50], cart=None),  # FBnoise.jl:24
}

# named configurations of BASELINE.json:configs
CONFIGS = {
    "P1": dict(system="P1", n=256, trials=1, theta_key="P1_MAX256", config_id=1),
    "P2": dict(system="P2", n=1000, trials=1, theta_key="P2_MAX1024", config_id=2),
    "CP": dict(system="CP", n=2000, trials=100, theta_key="CP_MAX2048", config_id=3),
    "FB": dict(system="FB", n=2000, trials=100, theta_key=None, config_id=4),
}

_HERE = os.path.dirname(os.path.abspath(__file__))


def _body_state(phi, psi, offset):
    """(13, m) block of one pendulum-like body at angles phi (m,), rates psi (m,)."""
    phi = np.asarray(phi, dtype=np.float64)
    s = np.zeros((13, phi.size))
    s[1] = 0.5 * np.sin(phi) + offset
    s[2] = -0.5 * np.cos(phi)
    s[3] = np.cos(phi / 2)
    s[4] = np.sin(phi / 2)
    s[8] = 0.5 * np.cos(phi) * psi
    s[9] = 0.5 * np.sin(phi) * psi
    s[10] = psi
    return s


def _cart_state(u, udot):
    u = np.asarray(u, dtype=np.float64)
    s = np.zeros((13, u.size))
    s[1] = u
    s[3] = 1.0
    s[8] = udot
    return s


def _step(phi, psi):
    psi2 = psi + DT * (-1.5 * 9.81 * np.sin(phi))
    return phi + DT * psi2, psi2


def make_trial(system: str, n: int, seed: int, noise: float = 1e-3, n_test: int = 0):
    """One trial's training set.  Returns dict X (d, n), Y (G, n) [+ Xtest (d, n_test)], Julia orientation."""
    spec = SYSTEMS[system]
    N = spec["bodies"]
    d = 13 * N
    rng = np.random.default_rng(seed)
    tot = n + n_test
    X = np.zeros((d, tot))
    Xn = np.zeros((d, tot))
    for b in range(N):
        if spec["cart"] == b:
            u = rng.uniform(-1, 1, tot)
            ud = rng.uniform(-1, 1, tot)
            force = rng.uniform(-1, 1, tot)
            X[13 * b:13 * b + 13] = _cart_state(u, ud)
            ud2 = ud + DT * force
            Xn[13 * b:13 * b + 13] = _cart_state(u + DT * ud2, ud2)
        else:
            phi = rng.uniform(-np.pi, np.pi, tot)
            psi = rng.uniform(-1, 1, tot) * 3.0
            X[13 * b:13 * b + 13] = _body_state(phi, psi, 0.25 * b)
            p2, s2 = _step(phi, psi)
            Xn[13 * b:13 * b + 13] = _body_state(p2, s2, 0.25 * b)
    varying = np.abs(X).max(axis=1) > 0
    varying &= X.std(axis=1) > 0
    X = X + noise * rng.standard_normal(X.shape) * varying[:, None]  # noise only on the non-constant rows
    idx = np.asarray(spec["outputs"]) - 1
    Y = Xn[idx, :] + noise * rng.standard_normal((idx.size, tot))
    out = {"X": np.asfortranarray(X[:, :n]), "Y": np.ascontiguousarray(Y[:, :n]), "system": system}
    if n_test:
        out["Xtest"] = np.asfortranarray(X[:, n:])
        out["Ytest"] = np.ascontiguousarray(Y[:, n:])
    return out


def load_theta_table():
    with open(os.path.join(_HERE, "theta0_config.json")) as f:
        return json.load(f)


def theta0(system: str, X: np.ndarray, key: str | None = None, log_noise: float = -2.0):
    """theta_0 = [logNoise, log l_1..l_d, log s_f] in GaussianProcesses order.

    With ``key`` the values come from the reference's examples/config/config.json entry of that name
    (``[s_f, l_1..l_d]`` natural units -> ``SEArd(log.(l), log(s_f))``, CPnoise.jl:38); the handful of keys the
    configs use are committed next to this file (theta0_config.json, extracted by tests/golden/make_theta0_config.py).  Otherwise the rule of
    examples/maximal_coordinates/FBparam.jl:23-26 without its random factor: s_f = 1, l_d = 10 / std_d (std 0 -> 1000)."""
    d = X.shape[0]
    if key is not None:
        v = np.asarray(load_theta_table()[key], dtype=np.float64)
        assert v.size == d + 1, (key, v.size, d)
        return np.concatenate([[log_noise], np.log(v[1:]), [np.log(v[0])]])
    std = X.std(axis=1)
    ell = np.where(std > 0, 10.0 / np.where(std > 0, std, 1.0), 1000.0)
    return np.concatenate([[log_noise], np.log(ell), [0.0]])


def make_config(name: str, trials: int | None = None, n: int | None = None, first_trial: int = 0, n_test: int = 0):
    """Workload of a named BASELINE config: list of trials, each {X, Y, theta0 (G, P)}; seed = 1000*config_id + trial."""
    cfg = CONFIGS[name]
    T = cfg["trials"] if trials is None else trials
    n = cfg["n"] if n is None else n
    out = []
    for t in range(first_trial, first_trial + T):
        tr = make_trial(cfg["system"], n, 1000 * cfg["config_id"] + t, n_test=n_test)
        th = theta0(cfg["system"], tr["X"], cfg["theta_key"])
        tr["theta0"] = np.tile(th, (tr["Y"].shape[0], 1))
        tr["trial"] = t
        out.append(tr)
    return out


def perturbed_thetas(theta0_, count: int, seed: int, scale: float = 0.1):
    """theta_0 plus seeded perturbations theta_0 + scale*N(0, I): distinct evaluation points for throughput runs."""
    rng = np.random.default_rng(seed)
    th = np.asarray(theta0_, dtype=np.float64)
    return [th] + [th + scale * rng.standard_normal(th.shape) for _ in range(count)]
