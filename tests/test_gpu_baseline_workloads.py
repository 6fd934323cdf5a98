"""GPU parity on the EXACT workloads BASELINE.json names and bench.py times (VERDICT r01 "parity gaps first"):
the CUDA path through the C ABI against the CPU oracle at the benchmark's own sizes, hyper-parameters and
perturbations, plus the make_posdef! retry ladder (info = k for k >= 2, info = -1).

Tolerances: kernel matrices 1e-12 relative; logML / gradient 1e-8 relative on well-conditioned inputs and
max(1e-8, 50 cond(K) eps) on the config.json hyper-parameters (s_f ~ 300-450, cond up to 1e10: two correct fp64
factorisations differ by ~cond*eps); predictions 1e-9 (same conditioning-aware widening)."""
import numpy as np
import pytest

from oracle import gp_oracle as go
from oracle.lbfgs_oracle import LBFGSOptions, lbfgs

pytestmark = pytest.mark.gpu
EPS = 2.220446049250313e-16


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def _batch(gprb, trials, thetas):
    gps = []
    for tr, th in zip(trials, thetas):
        for k in range(tr["Y"].shape[0]):
            t = th[k]
            gps.append(gprb.GPE(tr["X"], tr["Y"][k], gprb.MeanZero(), gprb.SEArd(t[1:-1], t[-1]), logNoise=t[0]))
    return gprb.GPBatch(gps)


def _check_eval(batch, b, mll, grad, info, r, what):
    cond = go.cond_estimate(r["state"]["K"])
    tol = max(1e-8, 50 * cond * EPS)
    assert info[b] == r["info"], (what, b, info[b], r["info"])
    assert abs(mll[b] - r["mll"]) <= tol * abs(r["mll"]), (what, b, cond, mll[b], r["mll"])
    assert rel(grad[b], r["grad"]) <= tol, (what, b, cond, rel(grad[b], r["grad"]))
    return cond, tol


def _check_optimum(r, o, f, th0, X, y):
    """Compare a batched-optimiser result with the scalar restatement, independent of the trajectory.  The searches drive
    logNoise to about -6.6 on these workloads: cond(K) ~ 1e9 at the later iterates, where two correct fp64 evaluations
    differ by ~cond*eps in value and gradient and ten to fifteen L-BFGS iterations amplify that.
    So: (1) the reported minimum is a CORRECT evaluation at the GPU's own minimiser - oracle at that point,
    conditioning-aware bound; (2) both searches improved on the start point, to the same objective level within 1 %."""
    xg = r["minimizer"]
    assert np.all(np.isfinite(xg)) and np.max(np.abs(xg)) < 300.0, xg
    rg = go.eval_mll(X, y, xg, with_grad=False, return_state=True)
    cond = go.cond_estimate(rg["state"]["K"])
    tol = max(1e-8, 50 * cond * EPS)
    assert rg["info"] == r["info"]
    assert abs(r["minimum"] + rg["mll"]) <= tol * abs(rg["mll"]), (cond, r["minimum"], -rg["mll"])
    f0 = f(th0)
    assert r["minimum"] < f0 and o.f < f0
    assert abs((f0 - r["minimum"]) / (f0 - o.f) - 1.0) <= 1e-2, (f0, r["minimum"], o.f)
    if cond * EPS < 1e-9:  # well-conditioned end point: the trajectories must agree closely
        assert abs(r["minimum"] - o.f) <= 1e-6 * abs(o.f)


def test_cp_n2000_bench_thetas(gprb):
    """configs[2] exactly as bench.py evaluates it: CP, n=2000, d=26, theta = config.json CP_MAX2048 + the bench's
    seeded 0.1 N(0,I) perturbations (rank 0, first timed step), the 8 GPs of the first two trials."""
    from gpr_jl_b200 import data
    trials = data.make_config("CP", trials=2)
    assert trials[0]["X"].shape == (26, 2000)
    theta0 = np.concatenate([tr["theta0"] for tr in trials])           # (8, 28)
    # bench.py: thetas = perturbed_thetas(batch.get_params(), steps + warmup, seed=1234 + rank) over ALL 400 GPs; the
    # first 8 rows of a (400, 28) normal draw are reproduced by drawing the full block
    full0 = np.tile(theta0[:1], (400, 1))
    pert = data.perturbed_thetas(full0, 3, seed=1234)[3][:8] - full0[:8]  # thetas[warmup + 0] with the default warmup = 3
    theta = theta0 + pert
    batch = _batch(gprb, trials, [theta[:4], theta[4:]])
    mll, grad, info = batch.eval(theta=theta, grad=True)
    hist = {}
    for b in range(8):
        tr = trials[b // 4]
        X = np.ascontiguousarray(tr["X"].T)
        r = go.eval_mll(X, tr["Y"][b % 4], theta[b], with_grad=True, return_state=True)
        _check_eval(batch, b, mll, grad, info, r, "CP n=2000")
        hist[int(info[b])] = hist.get(int(info[b]), 0) + 1
        if b in (0, 5):
            assert rel(batch.K(b), r["state"]["K"]) <= 1e-12
    assert sum(hist.values()) == 8
    batch.close()


def test_fb_d52_n2000(gprb):
    """configs[3]: FB fourbar, d=52, n=2000, rule-based theta_0 (config.json has no FB_MAX2048 entry) + perturbation."""
    from gpr_jl_b200 import data
    tr = data.make_config("FB", trials=1)[0]
    assert tr["X"].shape == (52, 2000) and tr["Y"].shape[0] == 12
    sel = [0, 7]
    rng = np.random.default_rng(52)
    theta = tr["theta0"][sel] + 0.1 * rng.standard_normal((2, 54))
    sub = {"X": tr["X"], "Y": tr["Y"][sel]}
    batch = _batch(gprb, [sub], [theta])
    mll, grad, info = batch.eval(theta=theta, grad=True)
    X = np.ascontiguousarray(tr["X"].T)
    Xs = data.make_trial("FB", 8, seed=5, n_test=100)["Xtest"]
    mu, var = batch.predict_y(Xs)
    for b in range(2):
        r = go.eval_mll(X, sub["Y"][b], theta[b], with_grad=True, return_state=True)
        cond, tol = _check_eval(batch, b, mll, grad, info, r, "FB n=2000")
        if b == 0:
            assert rel(batch.K(b), r["state"]["K"]) <= 1e-12
        m_o, v_o = go.predict(X, theta[b], r["state"], np.ascontiguousarray(Xs.T))
        ptol = max(1e-9, 50 * cond * EPS)
        assert rel(mu[b], m_o) <= ptol
        np.testing.assert_allclose(var[b], v_o, rtol=ptol, atol=1e-13)
    batch.close()


def test_p2_n1000_eval_and_optimize(gprb):
    """configs[1]: P2 double pendulum, d=26, n=1000, "hyperparameter optimisation per output dim": the evaluation at
    config.json P2_MAX1024 for all 6 outputs, and a 10-iteration optimize! of two outputs against the scalar Optim
    restatement driving the CPU oracle."""
    from gpr_jl_b200 import data
    tr = data.make_config("P2", trials=1)[0]
    assert tr["X"].shape == (26, 1000) and tr["Y"].shape[0] == 6
    theta = tr["theta0"]
    batch = _batch(gprb, [tr], [theta])
    mll, grad, info = batch.eval(theta=theta, grad=True)
    X = np.ascontiguousarray(tr["X"].T)
    for b in range(6):
        r = go.eval_mll(X, tr["Y"][b], theta[b], with_grad=True, return_state=True)
        _check_eval(batch, b, mll, grad, info, r, "P2 n=1000")
    batch.close()
    # optimisation: rule-based start (the search of hyperparameter.jl starts from such points, P2param.jl:24-27), 10 L-BFGS
    # iterations, outputs 0 and 3.  The targets get 0.05 N(0,1) of extra observation noise: with the generator's 1e-3 the
    # search drives logNoise below -6, cond(K) above 1e9, and two correct fp64 searches part ways after a few iterations
    #  - with a noise level the optimiser can find, the trajectories are comparable step for step.
    th0 = data.theta0("P2", tr["X"])
    sel = [0, 3]
    sub = {"X": tr["X"], "Y": tr["Y"][sel] + 0.05 * np.random.default_rng(1000).standard_normal((2, 1000))}
    ob = _batch(gprb, [sub], [np.tile(th0, (2, 1))])
    res = ob.optimize(gprb.LBFGS(linesearch=gprb.BackTracking(order=2)), gprb.Options(iterations=10))
    for k in range(2):
        y = sub["Y"][k]
        f = lambda t: -go.eval_mll(X, y, t, with_grad=False)["mll"]

        def fg(t):
            r = go.eval_mll(X, y, t)
            return (-r["mll"], -r["grad"]) if r["info"] >= 0 else (np.inf, np.full(t.size, np.nan))
        o = lbfgs(f, fg, th0, LBFGSOptions(iterations=10))
        assert res[k]["iterations"] == o.iterations
        _check_optimum(res[k], o, f, th0, X, y)
    ob.close()


def test_p1_n256_config_theta_and_optimize(gprb):
    """configs[0]: P1 simple pendulum, n=256, one trial (3 GPs), theta_0 = config.json P1_MAX256 (the reference's own
    CPU-runnable case): evaluation, prediction and a 15-iteration optimize! against the oracle."""
    from gpr_jl_b200 import data
    tr = data.make_config("P1", trials=1, n_test=20)[0]
    assert tr["X"].shape == (13, 256)
    theta = tr["theta0"]
    batch = _batch(gprb, [tr], [theta])
    mll, grad, info = batch.eval(theta=theta, grad=True)
    mu, var = batch.predict_y(tr["Xtest"])
    X = np.ascontiguousarray(tr["X"].T)
    for b in range(3):
        r = go.eval_mll(X, tr["Y"][b], theta[b], with_grad=True, return_state=True)
        cond, tol = _check_eval(batch, b, mll, grad, info, r, "P1 n=256")
        assert rel(batch.K(b), r["state"]["K"]) <= 1e-12
        m_o, v_o = go.predict(X, theta[b], r["state"], np.ascontiguousarray(tr["Xtest"].T))
        ptol = max(1e-9, 50 * cond * EPS)
        assert rel(mu[b], m_o) <= ptol
        np.testing.assert_allclose(var[b], v_o, rtol=ptol, atol=1e-12)
    th0 = data.theta0("P1", tr["X"])
    trn = {"X": tr["X"], "Y": tr["Y"] + 0.05 * np.random.default_rng(256).standard_normal(tr["Y"].shape)}  # see the P2 test
    ob = _batch(gprb, [trn], [np.tile(th0, (3, 1))])
    res = ob.optimize(gprb.LBFGS(linesearch=gprb.BackTracking(order=2)), gprb.Options(iterations=15))
    for k in range(3):
        y = trn["Y"][k]
        f = lambda t: -go.eval_mll(X, y, t, with_grad=False)["mll"]

        def fg(t):
            r = go.eval_mll(X, y, t)
            return (-r["mll"], -r["grad"]) if r["info"] >= 0 else (np.inf, np.full(t.size, np.nan))
        o = lbfgs(f, fg, th0, LBFGSOptions(iterations=15))
        assert res[k]["iterations"] == o.iterations
        _check_optimum(res[k], o, f, th0, X, y)


def test_sweep_point_n4096(gprb):
    """configs[4] beyond n=2048: one n=4096, d=26 GP of the sweep's input distribution (iid N(0,1), rule-based theta):
    K, logML, gradient and the m=100 prediction against the oracle."""
    n, d = 4096, 26
    rng = np.random.default_rng(4096)
    X = np.asfortranarray(rng.standard_normal((d, n)))
    y = np.sin(X[0]) + 0.1 * rng.standard_normal(n)
    th = np.concatenate([[-2.0], np.full(d, np.log(10.0)), [0.0]]) + 0.05 * rng.standard_normal(d + 2)
    batch = _batch(gprb, [{"X": X, "Y": y[None, :]}], [th[None, :]])
    mll, grad, info = batch.eval(grad=True)
    Xc = np.ascontiguousarray(X.T)
    r = go.eval_mll(Xc, y, th, with_grad=True, return_state=True)
    cond, tol = _check_eval(batch, 0, mll, grad, info, r, "n=4096")
    K = batch.K(0)
    assert rel(K, r["state"]["K"]) <= 1e-12
    del K
    Xs = np.asfortranarray(rng.standard_normal((d, 100)))
    mu, var = batch.predict_y(Xs)
    m_o, v_o = go.predict(Xc, th, r["state"], np.ascontiguousarray(Xs.T))
    ptol = max(1e-9, 50 * cond * EPS)
    assert rel(mu[0], m_o) <= ptol
    np.testing.assert_allclose(var[0], v_o, rtol=ptol, atol=1e-13)
    batch.close()


@pytest.mark.parametrize("n", [96, 300])
def test_make_posdef_ladder_info_k_and_minus_one(gprb, n):
    """make_posdef! adds 1e-6 tr(K)/n to the stored diagonal, cumulatively, up to 10 times.  A positive-semidefinite
    kernel never needs more than one addition by rounding alone, so the matrices that need exactly k are built with the
    fixed diagonal offset (gprb_batch_set_diag_offset): K - s I with s chosen against the smallest eigenvalue so that
    k = 0, 1, 2, 3, 5, 10 additions are needed, and one that is still indefinite after 10 (info = -1, mll = -Inf).
    Checked against oracle.chol_upper_jitter on the identical matrix."""
    from gpr_jl_b200 import data
    tr = data.make_trial("P1", n, seed=11 + n)
    th = data.theta0("P1", tr["X"])
    th[1:-1] -= 1.0
    X = np.ascontiguousarray(tr["X"].T)
    K = go.assemble_K(X, th)
    lam = float(np.linalg.eigvalsh(K)[0])
    T = float(np.trace(K)) / n
    want = [0, 1, 2, 3, 5, 10, -1]

    def shift_for(k):
        # offset -s makes the smallest eigenvalue lam - s + j delta after j additions, delta = 1e-6 (T - s) (the increment
        # is computed from the shifted matrix; later ones grow by ~1e-6 relative): exactly k additions are needed when
        # s = lam + (k - 1/2) delta
        delta = 1e-6 * (T - lam) / (1.0 + 1e-6 * (k - 0.5))
        return -(lam + (k - 0.5) * delta)
    shifts = [0.0] + [shift_for(k) for k in (1, 2, 3, 5, 10)] + [shift_for(12.5)]
    B = len(want)
    trB = {"X": tr["X"], "Y": np.tile(tr["Y"][:1], (B, 1))}
    thetas = np.tile(th, (B, 1))
    batch = _batch(gprb, [trB], [thetas])
    batch.set_diag_offset(np.array(shifts))
    mll, grad, info = batch.eval(theta=thetas, grad=True)
    for b in range(B):
        r = go.eval_mll(X, trB["Y"][b], th, with_grad=True, return_state=True, diag_offset=shifts[b])
        assert r["info"] == want[b], (b, r["info"], want[b])      # the construction does what it says, on LAPACK
        assert info[b] == want[b], (b, info[b], want[b])
        if want[b] < 0:
            assert mll[b] == -np.inf and np.all(np.isnan(grad[b]))
            continue
        cond = go.cond_estimate(r["state"]["K"])
        tol = max(1e-8, 50 * cond * EPS)
        assert abs(mll[b] - r["mll"]) <= tol * abs(r["mll"]), (b, cond)
        assert rel(grad[b], r["grad"]) <= tol, (b, cond)
        assert rel(np.diag(batch.K(b)), np.diag(r["state"]["K"])) <= 1e-12   # the jittered diagonal that was factorised
    # value-only passes walk the same ladder
    mll_v, _, info_v = batch.eval(theta=thetas, grad=False)
    assert np.array_equal(info_v, info) and np.array_equal(mll_v, mll)
    # the optimiser survives a GP that can never be factorised: it ends at once with info -1, the others optimise
    res = batch.optimize(gprb.LBFGS(linesearch=gprb.BackTracking(order=2)), gprb.Options(iterations=2))
    assert res[-1]["info"] == -1 and res[-1]["iterations"] == 0
    assert all(np.isfinite(r["minimum"]) for r in res[:-1])
    # ... and prediction masks it per GP instead of refusing the batch (examples/parallel/core.jl:41-46)
    mu, var = batch.predict_y(tr["X"][:, :7])
    assert np.all(np.isnan(mu[-1])) and np.all(np.isnan(var[-1]))
    assert np.all(np.isfinite(mu[:-1])) and np.all(var[:-1] > 0)
    batch.set_diag_offset(None)
    _, _, info0 = batch.eval(theta=thetas, grad=False)
    assert np.all(info0 == 0)
    batch.close()
