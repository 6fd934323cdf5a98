"""Scalar restatement of the optimiser the reference hands every GP to.  TEST INFRASTRUCTURE ONLY.

Reference call (identical in every experiment file, e.g.
examples/maximal_coordinates/CPnoise.jl:41):

    GaussianProcesses.optimize!(gp, LBFGS(linesearch = BackTracking(order=2)),
                                Optim.Options(time_limit=10.))

The algorithm lives in un-vendored packages pinned in /root/reference/Manifest.toml:
Optim 1.4.1 (:812-816), LineSearches 7.1.1 (:657-661), NLSolversBase 7.8.1
(:755-759).  **Parity unpinned** (no Julia here, no reference tests); this file
restates their published algorithm:

  * LBFGS(m=10, alphaguess=InitialStatic(alpha=1), scaleinvH0=true): two-loop
    recursion over a ring buffer indexed by ``pseudo_iteration``; H0 scaling
    gamma = s'y / y'y from the newest pair (identity on the first iteration);
    direction reset to -g when g's >= 0; history reset (pseudo_iteration=0)
    when 1/(dx'dg) is infinite.
  * BackTracking(c_1=1e-4, rho_hi=0.5, rho_lo=0.1, iterations=1000, order=2):
    function values only; "halve until finite" pre-loop (<= 52 halvings... the
    package uses ceil(-log2(eps)) = 52); quadratic interpolation clamped to
    [rho_lo, rho_hi] * alpha.
  * Optim.Options defaults: g_abstol=1e-8 on ||g||_inf, x/f tolerances 0
    (=> converged also when the step or the objective change is exactly 0),
    iterations=1000.  time_limit is wall-clock and therefore NOT reproducible;
    here stopping is by iteration / evaluation count (time_limit optional).
  * GaussianProcesses.get_optim_target: objective = -mll, gradient = -dmll,
    +Inf when the evaluation fails (PosDefException / non-finite theta).
  * After the accepted step the gradient is obtained with a second, full
    value+gradient evaluation at the new point (NLSolversBase value_gradient!).
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import numpy as np


@dataclass
class LBFGSOptions:
    m: int = 10
    g_abstol: float = 1e-8
    iterations: int = 1000
    max_evals: int = 0  # 0 = unlimited (value-only + value/grad evaluations)
    time_limit: float = math.inf
    # > 0: deterministic virtual clock instead of wall time - every value-only evaluation costs cost_value seconds and
    # every value+gradient evaluation cost_grad seconds (Optim checks time_limit once per iteration)
    cost_value: float = 0.0
    cost_grad: float = 0.0
    c_1: float = 1e-4
    rho_hi: float = 0.5
    rho_lo: float = 0.1
    ls_iterations: int = 1000
    iterfinite_max: int = 52  # ceil(-log2(eps(Float64)))


@dataclass
class LBFGSResult:
    x: np.ndarray
    f: float
    g_norm: float
    iterations: int
    f_calls: int
    fg_calls: int
    converged: bool
    ls_failed: bool
    stopped_by: str
    trace: list = field(default_factory=list)


def backtracking(phi, alpha0, phi_0, dphi_0, opt: LBFGSOptions):
    """LineSearches.BackTracking(order=2).  ``phi(alpha)`` -> objective value.
    Returns (alpha, phi_alpha, n_evals, ok)."""
    nev = 0
    a1 = a2 = alpha0
    phi1 = phi(a1)
    nev += 1
    iterfinite = 0
    while not math.isfinite(phi1) and iterfinite < opt.iterfinite_max:
        iterfinite += 1
        a1 = a2
        a2 = a1 / 2.0
        phi1 = phi(a2)
        nev += 1
    it = 0
    while phi1 > phi_0 + opt.c_1 * a2 * dphi_0:
        it += 1
        if it > opt.ls_iterations:
            return a2, phi1, nev, False
        # order 2 (and first iteration of order 3): quadratic interpolation
        denom = 2.0 * (phi1 - phi_0 - dphi_0 * a2)
        a_tmp = -(dphi_0 * a2 * a2) / denom if denom != 0.0 else math.nan
        a1 = a2
        # NaNMath.min / NaNMath.max: NaN-ignoring
        hi = a2 * opt.rho_hi
        a_tmp = hi if math.isnan(a_tmp) else min(a_tmp, hi)
        lo = a2 * opt.rho_lo
        a2 = lo if math.isnan(a_tmp) else max(a_tmp, lo)
        phi1 = phi(a2)
        nev += 1
    return a2, phi1, nev, True


def twoloop(g, rho, dx_hist, dg_hist, m, pseudo_iteration):
    """Optim.twoloop! with scaleinvH0=true and P=nothing.  Returns s = -H g."""
    lower = pseudo_iteration - m
    upper = pseudo_iteration - 1
    q = g.copy()
    alpha = np.zeros(m)
    for index in range(upper, lower - 1, -1):
        if index < 1:
            continue
        i = (index - 1) % m
        alpha[i] = rho[i] * float(dx_hist[i] @ q)
        q -= alpha[i] * dg_hist[i]
    if pseudo_iteration > 1:
        i = (upper - 1) % m
        scaling = float(dx_hist[i] @ dg_hist[i]) / float(dg_hist[i] @ dg_hist[i])
        s = scaling * q
    else:
        s = q.copy()
    for index in range(lower, upper + 1):
        if index < 1:
            continue
        i = (index - 1) % m
        beta = rho[i] * float(dg_hist[i] @ s)
        s += dx_hist[i] * (alpha[i] - beta)
    return -s


def lbfgs(f, fg, x0, opt: LBFGSOptions | None = None, keep_trace=False) -> LBFGSResult:
    """Minimise with Optim.LBFGS semantics.  f(x)->float ; fg(x)->(float, grad)."""
    opt = opt or LBFGSOptions()
    t0 = time.time()
    x = np.array(x0, dtype=np.float64, copy=True)
    n = x.size
    m = opt.m
    dx_hist = np.zeros((m, n))
    dg_hist = np.zeros((m, n))
    rho = np.zeros(m)
    pseudo_iteration = 0
    f_calls = 0
    fg_calls = 0
    fx, g = fg(x)
    fg_calls += 1
    g = np.array(g, dtype=np.float64, copy=True)
    trace = []
    res = dict(converged=False, ls_failed=False, stopped_by="iterations")

    def g_res(gv):
        return float(np.max(np.abs(gv))) if gv.size else 0.0

    # initial_convergence
    if not math.isfinite(fx) or not np.all(np.isfinite(g)):
        return LBFGSResult(x, fx, math.nan, 0, f_calls, fg_calls, False, False, "nonfinite_start", trace)
    if g_res(g) <= opt.g_abstol:
        return LBFGSResult(x, fx, g_res(g), 0, f_calls, fg_calls, True, False, "g_abstol", trace)

    vtime = opt.cost_value > 0.0 or opt.cost_grad > 0.0
    iteration = 0
    while iteration < opt.iterations:
        iteration += 1
        # ---- update_state!
        pseudo_iteration += 1
        s = twoloop(g, rho, dx_hist, dg_hist, m, pseudo_iteration)
        g_prev = g.copy()
        dphi_0 = float(g @ s)
        if dphi_0 >= 0.0:  # reset_search_direction!
            pseudo_iteration = 1
            s = -g
            dphi_0 = float(g @ s)
        phi_0 = fx

        def phi(a):
            return f(x + a * s)

        alpha, _, nev, ok = backtracking(phi, 1.0, phi_0, dphi_0, opt)
        f_calls += nev
        dx = alpha * s
        x_prev = x
        f_prev = fx
        x = x + dx
        if not ok:
            res.update(ls_failed=True, stopped_by="linesearch")
            break
        # ---- update_g!  (value_gradient! at the new point: a second full evaluation)
        fx_new, g_new = fg(x)
        fg_calls += 1
        g_new = np.array(g_new, dtype=np.float64, copy=True)
        if keep_trace:
            trace.append((iteration, fx_new, g_res(g_new), alpha, nev))
        # ---- assess_convergence (x/f tolerances are 0 => only exact stalls count)
        x_conv = float(np.max(np.abs(x - x_prev))) <= 0.0
        f_conv = abs(fx_new - f_prev) <= 0.0
        g_conv = g_res(g_new) <= opt.g_abstol  # maximum(abs, g): NaN never converges
        # f_increased (f_x > f_x_previous; +Inf when the re-evaluation of the accepted point failed) stops the run
        # (allow_f_increases = false) and pick_best_x returns the PREVIOUS point; a non-finite gradient terminates
        # too ("Terminated early due to NaN in gradient").  Either way the last good x, f, g are kept.
        if fx_new > f_prev or not math.isfinite(fx_new) or not np.all(np.isfinite(g_new)):
            x = x_prev
            res.update(converged=bool(x_conv or f_conv or g_conv), stopped_by="f_increased")
            break
        fx, g = fx_new, g_new
        if x_conv or f_conv or g_conv:
            res.update(converged=True, stopped_by="g_abstol" if g_conv else ("x_stall" if x_conv else "f_stall"))
            break
        # ---- update_h!
        dg = g - g_prev
        denom = float(dx @ dg)
        rho_it = math.inf if denom == 0.0 else 1.0 / denom
        if math.isinf(rho_it):
            pseudo_iteration = 0
        else:
            idx = (pseudo_iteration - 1) % m
            dx_hist[idx] = dx
            dg_hist[idx] = dg
            rho[idx] = rho_it
        if opt.max_evals and f_calls + fg_calls >= opt.max_evals:
            res.update(stopped_by="max_evals")
            break
        if vtime:
            if f_calls * opt.cost_value + fg_calls * opt.cost_grad > opt.time_limit:
                res.update(stopped_by="time_limit")
                break
        elif time.time() - t0 > opt.time_limit:
            res.update(stopped_by="time_limit")
            break
    return LBFGSResult(x, fx, g_res(g), iteration, f_calls, fg_calls, res["converged"], res["ls_failed"], res["stopped_by"], trace)
