"""Batched experiment body: the GP part of the reference's ``_experiment_<sys>_max`` functions
(/root/reference/examples/maximal_coordinates/CPnoise.jl:12-54 and siblings) for many trials at once.

Per trial the reference does, sequentially on one CPU thread:
    for each output k:  kernel = SEArd(log.(params[2:end]), log(params[1])); gp = GP(X, y_k, mean, kernel); optimize!(gp, ...)
    for each test state: predictdynamics(mechanism, gps, x0, steps, getvomega)      (examples/utils/predictdynamics.jl:7-22)
Here all G x T GPs of the trials resident on this GPU are built into ONE GPBatch, optimised in lock-step with one
library call, and every rollout step predicts all (trial, test state, GP) means in one call.  The physics between
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


def predictdynamics(batch, trials_G, start_states, steps, step_fn, var=False):
    """Batched ``predictdynamics`` (examples/utils/predictdynamics.jl:7-22).

    batch        GPBatch holding the GPs of T trials, trial-major (G consecutive GPs per trial)
    trials_G     G (GPs per trial)
    start_states list over trials of (d, m) arrays - the m test states of each trial
    step_fn(trial, states (d,m), mu (G,m)) -> next states (d,m): host callback doing getvomega + projectv! + updatestate!
    Returns the list of final (d, m) state blocks.  One gprb_predict call per rollout step for all trials/states/GPs."""
    T = len(start_states)
    assert batch.B == T * trials_G
    states = [np.asfortranarray(np.asarray(s, dtype=np.float64)) for s in start_states]
    for _ in range(steps):
        blocks = [states[b // trials_G] for b in range(batch.B)]
        mu, _ = batch.predict_y(blocks, var=var, per_gp=True)
        for t in range(T):
            states[t] = np.asfortranarray(step_fn(t, states[t], mu[t * trials_G:(t + 1) * trials_G]))
    return states
