"""The batched lock-step optimiser in libgprb200 (csrc/lbfgs.cu) against the scalar restatement of
Optim.LBFGS + LineSearches.BackTracking(order=2) (oracle/lbfgs_oracle.py) on an analytic objective.
Runs without a GPU: gprb_lbfgs_selftest evaluates B Rosenbrock copies on the host through the same state machine
gprb_optimize drives with gprb_eval."""
import ctypes as C

import numpy as np
import pytest

from oracle.lbfgs_oracle import LBFGSOptions, lbfgs
import gpr_jl_b200  # noqa: F401
from gpr_jl_b200.lib import LbfgsOpts, OptResult, load_library, _d


def rosen(x):
    """Same operation order as the C self-test objective (bitwise-identical doubles)."""
    s = 0.0
    for p in range(len(x) - 1):
        t1 = x[p + 1] - x[p] * x[p]
        t2 = 1.0 - x[p]
        s += 100.0 * t1 * t1 + t2 * t2
    return s


def rosen_der(x):
    g = np.zeros(len(x))
    for p in range(len(x) - 1):
        t1 = x[p + 1] - x[p] * x[p]
        t2 = 1.0 - x[p]
        g[p] += -400.0 * x[p] * t1 - 2.0 * t2
        g[p + 1] += 200.0 * t1
    return g


def _f(bound):
    def f(x):
        x = [float(v) for v in x]
        return rosen(x) if all(abs(v) <= bound for v in x) else np.inf

    def fg(x):
        x = [float(v) for v in x]
        if all(abs(v) <= bound for v in x):
            return rosen(x), rosen_der(x)
        return np.inf, np.full(len(x), np.nan)
    return f, fg


def _run_lib(x0, bound, **kw):
    lib = load_library()
    B, P = x0.shape
    o = LbfgsOpts()
    lib.dll.gprb_lbfgs_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    th = np.ascontiguousarray(x0, dtype=np.float64).copy()
    res = (OptResult * B)()
    lib.check(lib.dll.gprb_lbfgs_selftest(B, P, _d(th), C.byref(o), bound, res))
    return th, res


@pytest.mark.parametrize("bound", [1e9, 2.5])
def test_batched_trajectories_match_scalar_oracle(bound):
    rng = np.random.default_rng(3)
    B, P = 9, 6
    x0 = rng.uniform(-2, 2, (B, P))
    th, res = _run_lib(x0, bound)
    f, fg = _f(bound)
    for b in range(B):
        r = lbfgs(f, fg, x0[b], LBFGSOptions())
        assert res[b].iterations == r.iterations, (b, res[b].iterations, r.iterations)
        assert res[b].f_calls == r.f_calls and res[b].fg_calls == r.fg_calls
        assert bool(res[b].converged) == r.converged
        np.testing.assert_allclose(th[b], r.x, rtol=1e-12, atol=1e-12)
        assert res[b].mll == pytest.approx(r.f, rel=1e-12, abs=1e-20)


def test_iteration_and_eval_caps_are_per_gp():
    rng = np.random.default_rng(4)
    x0 = rng.uniform(-2, 2, (5, 4))
    th, res = _run_lib(x0, 1e9, iterations=7)
    f, fg = _f(1e9)
    for b in range(5):
        r = lbfgs(f, fg, x0[b], LBFGSOptions(iterations=7))
        assert res[b].iterations == r.iterations <= 7
        np.testing.assert_allclose(th[b], r.x, rtol=1e-12, atol=1e-12)
    th, res = _run_lib(x0, 1e9, max_evals=25)
    for b in range(5):
        r = lbfgs(f, fg, x0[b], LBFGSOptions(max_evals=25))
        assert res[b].f_calls + res[b].fg_calls == r.f_calls + r.fg_calls
        np.testing.assert_allclose(th[b], r.x, rtol=1e-12, atol=1e-12)


def test_oracle_reaches_the_rosenbrock_minimum():
    f, fg = _f(1e9)
    r = lbfgs(f, fg, np.array([-1.2, 1.0, -0.5, 0.8]))
    assert r.converged and r.g_norm <= 1e-8
    np.testing.assert_allclose(r.x, np.ones(4), atol=1e-6)


def test_start_at_optimum_and_nonfinite_start():
    th, res = _run_lib(np.ones((2, 3)), 1e9)
    assert res[0].iterations == 0 and res[0].converged == 1
    th, res = _run_lib(np.full((1, 3), 5.0), 2.0)  # outside the box: objective is +Inf at the start
    assert res[0].iterations == 0 and res[0].converged == 0


def test_per_gp_virtual_time_limit_matches_oracle():
    """time_limit with evaluation costs set runs on a deterministic per-GP virtual clock (the reference's
    Optim.Options(time_limit=10.) is 10 s of CPU time per GP, CPnoise.jl:41): every GP stops at the first iteration
    boundary where its own f_calls * cost_value + fg_calls * cost_grad exceeds the limit - GPs whose line searches
    backtrack more stop after fewer iterations, exactly like the scalar restatement."""
    rng = np.random.default_rng(11)
    x0 = rng.uniform(-2, 2, (12, 5))
    th, res = _run_lib(x0, 1e9, time_limit=10.0, cost_value=0.35, cost_grad=0.9)
    f, fg = _f(1e9)
    its = set()
    for b in range(12):
        r = lbfgs(f, fg, x0[b], LBFGSOptions(time_limit=10.0, cost_value=0.35, cost_grad=0.9))
        assert r.stopped_by == "time_limit"
        assert (res[b].iterations, res[b].f_calls, res[b].fg_calls) == (r.iterations, r.f_calls, r.fg_calls)
        assert res[b].f_calls * 0.35 + res[b].fg_calls * 0.9 > 10.0
        np.testing.assert_allclose(th[b], r.x, rtol=1e-12, atol=1e-12)
        its.add((res[b].iterations, res[b].f_calls))
    assert len(its) > 1  # the limit really is per GP: different GPs stop after different amounts of work


def test_gradient_inf_norm_propagates_nan():
    """maximum(abs, g) is NaN when any entry is NaN (Julia semantics): a failed gradient never counts as converged
    (fmax dropped the NaN and reported ||g|| = 0)."""
    # start inside the box but with the first step leaving it: the objective returns +Inf / NaN gradient outside
    th, res = _run_lib(np.array([[1.9, 1.9, 1.9]]), 2.0, iterations=3)
    assert np.all(np.isfinite(th)) and np.all(np.abs(th) <= 2.0)  # the optimiser never keeps a point outside the box
    assert np.isfinite(res[0].mll)
