"""Independent pin of the oracle's mathematics: scikit-learn's GaussianProcessRegressor (a separate, published
implementation of exact GP regression) against oracle/gp_oracle.py on identical inputs.

This does not replace a run of the reference (GaussianProcesses.jl is unavailable: "parity unpinned"),
but it removes the risk that the restated formulas of SURVEY.md A.1-A.3/A.6 are self-consistent yet wrong: log marginal
likelihood, its gradient in log-parameters, predictive mean and predict_y variance agree to ~1e-14 for the SEArd kernel the
reference uses and for the three Matern-ARD extensions.

Parameter maps:  sklearn kernel = ConstantKernel(s_f^2) * RBF|Matern(length_scale = l) + WhiteKernel(s_n^2),
sklearn theta = log [s_f^2, l_1..l_d, s_n^2]  =>  d/d lsigma = 2 d/d log s_f^2,  d/d logNoise = 2 d/d log s_n^2
(the reference's extra eps() on the noise diagonal is below the comparison tolerance)."""
import numpy as np
import pytest

from oracle import gp_oracle as go

sk = pytest.importorskip("sklearn.gaussian_process")
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel  # noqa: E402


def _sk_kernel(kind, ell, sf2, sn2):
    base = {"se": lambda: RBF(length_scale=ell), "mat12": lambda: Matern(length_scale=ell, nu=0.5),
            "mat32": lambda: Matern(length_scale=ell, nu=1.5), "mat52": lambda: Matern(length_scale=ell, nu=2.5)}[kind]()
    return ConstantKernel(sf2) * base + WhiteKernel(sn2)


@pytest.mark.parametrize("kind", ["se", "mat12", "mat32", "mat52"])
@pytest.mark.parametrize("system,n", [("CP", 120), ("P1", 64)])
def test_oracle_matches_scikit_learn(kind, system, n):
    import gpr_jl_b200  # noqa: F401
    from gpr_jl_b200 import data
    tr = data.make_trial(system, n, seed=3 + n, n_test=6)
    th = data.theta0(system, tr["X"])
    th[1:-1] -= 1.0
    th += 0.05 * np.random.default_rng(n).standard_normal(th.size)
    X, Xs, y = np.ascontiguousarray(tr["X"].T), np.ascontiguousarray(tr["Xtest"].T), tr["Y"][1]
    ell, sf2, sn2 = np.exp(th[1:-1]), np.exp(2 * th[-1]), np.exp(2 * th[0])
    gpr = sk.GaussianProcessRegressor(kernel=_sk_kernel(kind, ell, sf2, sn2), optimizer=None, alpha=0.0).fit(X, y)
    lml, g_sk = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
    r = go.eval_mll(X, y, th, kind=kind, with_grad=True, return_state=True)
    g_ours = np.concatenate([[r["grad"][-1] / 2], r["grad"][1:-1], [r["grad"][0] / 2]])  # sklearn order and log-variances
    assert abs(lml - r["mll"]) <= 1e-11 * abs(lml)
    assert np.max(np.abs(g_sk - g_ours)) <= 1e-10 * np.max(np.abs(g_sk))
    mu_sk, std_sk = gpr.predict(Xs, return_std=True)
    mu, var = go.predict(X, th, r["state"], Xs, kind=kind)
    assert np.max(np.abs(mu_sk - mu)) <= 1e-10 * np.max(np.abs(mu_sk))
    np.testing.assert_allclose(std_sk ** 2, var, rtol=1e-9)
