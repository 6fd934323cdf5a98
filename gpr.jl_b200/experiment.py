"""Batched experiment body: the GP part of the reference's ``_experiment_<sys>_max`` functions
(/root/reference/examples/maximal_coordinates/CPnoise.jl:12-54 and siblings) for many trials at once.

Per trial the reference does, sequentially on one CPU thread:
    for each output k:  kernel = SEArd(log.(params[2:end]), log(params[1])); gp = GP(X, y_k, mean, kernel); optimize!(gp, ...)
    for each test state: predictdynamics(mechanism, gps, x0, steps, getvomega)      (examples/utils/predictdynamics.jl:7-22)
Here all G x T GPs of the trials resident on this GPU are built into ONE GPBatch, optimised in lock-step with one
library call, and every rollout step predicts all (trial, test state, GP) means of a trial group in one call, two
groups alternating so that the device predicts one while the host projects the other.  The physics between
two steps - ``getvomega`` -> ``projectv!`` -> ``updatestate!`` - stays a host callback (ConstrainedDynamics.jl is not
part of this path).
"""
from __future__ import annotations

import numpy as np

from . import gp as _gp


def params_to_theta(params, log_noise=-2.0):
    """config.json order ``[s_f, l_1..l_d]`` -> GaussianProcesses order ``[logNoise, ll_1..ll_d, lsigma]`` (CPnoise.jl:38)."""
    params = np.asarray(params, dtype=np.float64)
    return np.concatenate([[log_noise], np.log(params[1:]), [np.log(params[0])]])


def build_batch(trials, params, mean_factory=None, kernel=_gp.SEArd, log_noise=-2.0):
    """trials: list of {X (d,n), Y (G,n)}; params: config.json-style vector shared by all GPs (as in the reference)
    or a list of per-trial (G, d+1) arrays.  mean_factory(trial_index, k) -> Mean (default MeanZero)."""
    gps = []
    for ti, tr in enumerate(trials):
        for k in range(tr["Y"].shape[0]):
            p = params if np.ndim(params) == 1 else params[ti][k]
            p = np.asarray(p, dtype=np.float64)
            mean = mean_factory(ti, k) if mean_factory else _gp.MeanZero()
            gps.append(_gp.GPE(tr["X"], tr["Y"][k], mean, kernel(np.log(p[1:]), float(np.log(p[0]))), logNoise=log_noise))
    return _gp.GPBatch(gps)


def fit_trials(trials, params, mean_factory=None, method=None, options=None):
    """GP(...) + optimize!(...) for every output of every trial, batched.  Returns (batch, per-GP optimiser results)."""
    batch = build_batch(trials, params, mean_factory)
    res = batch.optimize(method or _gp.LBFGS(linesearch=_gp.BackTracking(order=2)), options or _gp.Options())
    return batch, res


def predictdynamics(batch, trials_G, start_states, steps, step_fn, var=False, overlap=True, timing=None):
    """Batched ``predictdynamics`` (examples/utils/predictdynamics.jl:7-22).

    batch        GPBatch holding the GPs of T trials, trial-major (G consecutive GPs per trial)
    trials_G     G (GPs per trial)
    start_states list over trials of (d, m) arrays - the m test states of each trial
    step_fn(trial, states (d,m), mu (G,m)) -> next states (d,m): host callback doing getvomega + projectv! + updatestate!
                 (src/projections/implicitProjection.jl:80-107 stays on the host)
    Returns the list of final (d, m) state blocks.

    overlap=True (default, T >= 2): the trials are cut into two groups that alternate on the two prediction pipelines
    of the library (gprb_predict_async / gprb_predict_wait): while the host projects the states of one group, the device
    predicts the other - the per-step chain  H2D states -> predict -> D2H mu -> host projectv!  of predictdynamics.jl:11-19
    is hidden behind the other group's host work instead of serialising with it.  The G GPs of a trial share one
    uploaded state block.  overlap=False: one synchronous gprb_predict per step for all trials (same results, bit for bit).
    timing: optional dict that receives host_s (time inside step_fn), wait_s (host blocked on the device), total_s."""
    import time
    T = len(start_states)
    G = trials_G
    assert batch.B == T * G
    states = [np.asfortranarray(np.asarray(s, dtype=np.float64)) for s in start_states]
    t_host = t_wait = 0.0
    t_start = time.perf_counter()
    if not overlap or T < 2:
        for _ in range(steps):
            t0 = time.perf_counter()
            batch.predict_async(0, 0, batch.B, states, var=var, gps_per_block=G)
            mu, _ = batch.predict_wait(0)
            t1 = time.perf_counter()
            for t in range(T):
                states[t] = np.asfortranarray(step_fn(t, states[t], mu[t * G:(t + 1) * G]))
            t_wait += t1 - t0
            t_host += time.perf_counter() - t1
    else:
        cut = (0, T // 2, T)  # group g = trials cut[g] .. cut[g+1]-1 on pipeline slot g
        for g in (0, 1):
            batch.predict_async(g, cut[g] * G, cut[g + 1] * G, states[cut[g]:cut[g + 1]], var=var, gps_per_block=G)
        for step in range(steps):
            for g in (0, 1):
                t0 = time.perf_counter()
                mu, _ = batch.predict_wait(g)
                t1 = time.perf_counter()
                for t in range(cut[g], cut[g + 1]):
                    k = t - cut[g]
                    states[t] = np.asfortranarray(step_fn(t, states[t], mu[k * G:(k + 1) * G]))
                t2 = time.perf_counter()
                if step + 1 < steps:  # next step of this group goes to the device while the host turns to the other group
                    batch.predict_async(g, cut[g] * G, cut[g + 1] * G, states[cut[g]:cut[g + 1]], var=var, gps_per_block=G)
                t_wait += t1 - t0
                t_host += t2 - t1
    if timing is not None:
        timing.update(host_s=t_host, wait_s=t_wait, total_s=time.perf_counter() - t_start)
    return states
