"""CPU tests of the oracle itself (oracle/gp_oracle.py): the reference pins nothing on this path
(SURVEY.md section 4 - no tests, no golden vectors), so the restatement is anchored by finite differences, an
extended-precision arbiter, algebraic invariants, and the edge cases the reference's data exhibits."""
import numpy as np
import pytest

from oracle import gp_oracle as go
import gpr_jl_b200  # noqa: F401
from gpr_jl_b200 import data


def _case(system="P1", n=60, seed=7, rule=True):
    tr = data.make_trial(system, n, seed)
    X = np.ascontiguousarray(tr["X"].T)
    th = data.theta0(system, tr["X"]) if rule else data.theta0(system, tr["X"], "P1_MAX64")
    return X, tr["Y"][0], th


@pytest.mark.parametrize("kind", ["se", "mat12", "mat32", "mat52"])
def test_gradient_matches_central_differences(kind):
    X, y, th = _case(n=50)
    th = th.copy()
    th[1:-1] -= 1.5  # shorter length-scales so every dimension matters
    r = go.eval_mll(X, y, th, kind=kind)
    g = r["grad"]
    for p in range(th.size):
        h = 1e-5
        tp, tm = th.copy(), th.copy()
        tp[p] += h
        tm[p] -= h
        fd = (go.eval_mll(X, y, tp, kind=kind, with_grad=False)["mll"] - go.eval_mll(X, y, tm, kind=kind, with_grad=False)["mll"]) / (2 * h)
        assert abs(fd - g[p]) <= 2e-6 * max(1.0, abs(g[p])), (kind, p, fd, g[p])


def test_constant_dimensions_have_exactly_zero_gradient():
    X, y, th = _case(n=40)
    const = np.where(X.std(axis=0) == 0)[0]
    assert const.size >= 5  # planar systems: x_x, q_y, q_z, v_x, omega_y, omega_z are structurally constant
    g = go.eval_mll(X, y, th)["grad"]
    assert np.all(g[1 + const] == 0.0)


@pytest.mark.parametrize("kind", ["se", "mat52"])
def test_against_longdouble_arbiter(kind):
    X, y, th = _case(n=80)
    r = go.eval_mll(X, y, th, kind=kind, return_state=True)
    a = go.longdouble_eval(X, y, th, kind=kind)
    assert abs(r["mll"] - float(a["mll"])) <= 1e-10 * abs(float(a["mll"]))
    np.testing.assert_allclose(r["grad"], a["grad"].astype(np.float64), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(r["state"]["K"], a["K"].astype(np.float64), rtol=1e-13, atol=0)
    np.testing.assert_allclose(r["state"]["alpha"], a["alpha"].astype(np.float64), rtol=1e-8, atol=1e-10)


def test_factor_and_state_invariants():
    X, y, th = _case("P2", n=90)
    r = go.eval_mll(X, y, th, return_state=True)
    s = r["state"]
    U = np.triu(s["U"])
    np.testing.assert_allclose(U.T @ U, s["K"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(s["K"] @ s["alpha"], y, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(s["Kinv"] @ s["K"], np.eye(90), atol=1e-7)
    assert np.all(np.diag(s["K"]) == np.exp(2 * th[-1]) + (np.exp(2 * th[0]) + go.EPS))


def test_jitter_retry_and_failure_codes():
    rng = np.random.default_rng(0)
    # rank-deficient "kernel matrix": duplicate points, no noise -> not PD until jitter is added
    M = rng.standard_normal((30, 3))
    K = M @ M.T
    U, info, Kj = go.chol_upper_jitter(K)
    assert info >= 1 and U is not None
    assert np.allclose(np.triu(U).T @ np.triu(U), Kj)
    # hopeless matrix: negative definite
    U, info, _ = go.chol_upper_jitter(-np.eye(5))
    assert U is None and info == -1
    X, y, th = _case(n=20)
    bad = th.copy()
    bad[2] = np.nan
    r = go.eval_mll(X, y, bad)
    assert r["info"] == -2 and r["mll"] == -np.inf and np.all(np.isnan(r["grad"]))


def test_huge_lengthscales_like_cp_config():
    """config.json CP_MAX has l up to 3.5e37 on constant dims: weights underflow to ~0 without harm."""
    X, y, th = _case(n=40)
    th = th.copy()
    th[3] = np.log(3.5e37)
    r = go.eval_mll(X, y, th)
    assert np.isfinite(r["mll"]) and np.all(np.isfinite(r["grad"]))
    assert abs(r["grad"][3]) < 1e-60


def test_predict_matches_direct_formulas():
    X, y, th = _case(n=70)
    Xs = X[:5] + 0.01
    r = go.eval_mll(X, y, th, return_state=True)
    mu, var = go.predict(X, th, r["state"], Xs)
    K = r["state"]["K"]
    kc = go.cov_f(X, th, Xb=Xs)
    np.testing.assert_allclose(mu, kc.T @ np.linalg.solve(K, y), rtol=1e-9, atol=1e-12)
    v = np.exp(2 * th[-1]) - np.einsum("ij,ij->j", kc, np.linalg.solve(K, kc)) + np.exp(2 * th[0])
    np.testing.assert_allclose(var, v, rtol=1e-7, atol=1e-12)
    # at a training point the latent variance is small but the noise floor stays
    mu0, var0 = go.predict(X, th, r["state"], X[:1])
    assert var0[0] >= np.exp(2 * th[0])


def test_mean_offset_enters_only_through_targets():
    X, y, th = _case(n=30)
    m = np.linspace(-1, 1, 30)
    a = go.eval_mll(X, y - m, th)
    b = go.eval_mll(X, (y - m).copy(), th)
    assert a["mll"] == b["mll"]
