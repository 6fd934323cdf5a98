"""CPU oracle for the GPR.jl GP-regression hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``gpr.jl_b200``) never routes through it and has no CPU fallback.

**Parity status: UNPINNED.**  The arithmetic of this path does not live in
``/root/reference`` at all: the reference only *calls* the un-vendored Julia
packages pinned in ``/root/reference/Manifest.toml``

  * GaussianProcesses 0.12.4 (Manifest.toml:409-413)  - SEArd, GPE, update_mll!,
    update_mll_and_dmll!, predict_y
  * PDMats 0.10.1 (Manifest.toml:835-839)             - PDMat, ``\\``, logdet, whiten!
  * Optim 1.4.1 / LineSearches 7.1.1 (Manifest.toml:812-816, 657-661) - see lbfgs_oracle.py
  * stdlib LinearAlgebra -> OpenBLAS LAPACK (dpotrf, dpotrs, dtrtrs)

Julia is not installed in this image, the reference holds no tests, golden
vectors or seeds for this path, so nothing can pin this
restatement against reference outputs.  What this file does instead: it
restates the *published algorithm* of those package versions in the reference's
operation order, using the same LAPACK family (scipy -> OpenBLAS), and is itself
checked by finite differences and an extended-precision arbiter
(``longdouble_eval``) in ``tests/test_oracle.py``, and against scikit-learn's independent
GaussianProcessRegressor in ``tests/test_oracle_vs_sklearn.py`` (agreement ~1e-14).

Reference call sites that define the boundary this oracle restates:
  * examples/maximal_coordinates/CPnoise.jl:38-41  SEArd(log.(l), log(sf)); GP(X, y, mean, kernel); optimize!
  * examples/maximal_coordinates/{P1noise.jl:35-38, P2noise.jl:34-37, FBnoise.jl:33-36}
  * examples/utils/predictdynamics.jl:13           predict_y(gp, obs)[1][1]
  * src/mDynamics.jl:29-55                         zero-parameter Mean plug-in

Parameter vector (GaussianProcesses ``get_params(gp)`` order):
    theta = [logNoise, ll_1 .. ll_d, lsigma]        (length P = d + 2)
X is d x n column-major in Julia (one sample per column).  numpy arrays here are
shaped (n, d) C-contiguous, which is the *same memory*.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import lapack

EPS = float(np.finfo(np.float64).eps)  # Julia eps() added to the noise diagonal
LOG2PI = math.log(2.0 * math.pi)
MAX_JITTER = 10  # make_posdef! retries

KERNELS = {"se": 0, "mat12": 1, "mat32": 2, "mat52": 3}


# --------------------------------------------------------------------------
# covariance (GaussianProcesses kernels/se_ard.jl, stationary.jl; Matern per SURVEY A.3)
# --------------------------------------------------------------------------
def _wsqdist(Xa: np.ndarray, Xb: np.ndarray, w: np.ndarray) -> np.ndarray:
    """Weighted squared distance by direct differences (what Distances.pairwise!
    + the ARD dist_stack compute): r2_ij = sum_p w_p (Xa_ip - Xb_jp)^2,
    accumulated p = 1..d in order."""
    r2 = np.zeros((Xa.shape[0], Xb.shape[0]))
    for p in range(Xa.shape[1]):
        diff = Xa[:, p][:, None] - Xb[:, p][None, :]
        r2 += w[p] * (diff * diff)
    return r2


def _kfun(r2: np.ndarray, sf2: float, kind: str) -> np.ndarray:
    if kind == "se":
        return sf2 * np.exp(-0.5 * r2)
    r = np.sqrt(r2)
    if kind == "mat12":
        return sf2 * np.exp(-r)
    if kind == "mat32":
        s = math.sqrt(3.0) * r
        return sf2 * (1.0 + s) * np.exp(-s)
    if kind == "mat52":
        s = math.sqrt(5.0) * r
        return sf2 * (1.0 + s + 5.0 * r2 / 3.0) * np.exp(-s)
    raise ValueError(kind)


def _dk_dll_factor(r2: np.ndarray, sf2: float, kind: str) -> np.ndarray:
    """g(r) such that dk/dll_p = g(r) * w_p * Delta_p^2  (SURVEY A.2/A.3)."""
    if kind == "se":
        return sf2 * np.exp(-0.5 * r2)
    r = np.sqrt(r2)
    if kind == "mat12":
        with np.errstate(divide="ignore", invalid="ignore"):
            g = sf2 * np.exp(-r) / r
        return np.where(r > 0.0, g, 0.0)
    if kind == "mat32":
        return 3.0 * sf2 * np.exp(-math.sqrt(3.0) * r)
    if kind == "mat52":
        s = math.sqrt(5.0) * r
        return (5.0 / 3.0) * sf2 * (1.0 + s) * np.exp(-s)
    raise ValueError(kind)


def unpack_theta(theta: np.ndarray, d: int):
    theta = np.asarray(theta, dtype=np.float64)
    assert theta.shape == (d + 2,)
    log_noise = float(theta[0])
    w = np.exp(-2.0 * theta[1 : d + 1])  # SEArd stores il2 = exp.(-2ll)
    sf2 = math.exp(2.0 * float(theta[d + 1]))  # s2 = exp(2 lsigma)
    return log_noise, w, sf2


def cov_f(X: np.ndarray, theta: np.ndarray, kind: str = "se", Xb: np.ndarray | None = None) -> np.ndarray:
    """Noise-free kernel matrix K_f(X, Xb) (``cov(kernel, X, Xb)``)."""
    _, w, sf2 = unpack_theta(theta, X.shape[1])
    return _kfun(_wsqdist(X, X if Xb is None else Xb, w), sf2, kind)


def assemble_K(X: np.ndarray, theta: np.ndarray, kind: str = "se") -> np.ndarray:
    """``update_cK!``: K = K_f + (exp(2 logNoise) + eps()) I."""
    log_noise, _, _ = unpack_theta(theta, X.shape[1])
    K = cov_f(X, theta, kind)
    K[np.diag_indices_from(K)] += math.exp(2.0 * log_noise) + EPS
    return K


# --------------------------------------------------------------------------
# factorisation with the make_posdef! jitter loop
# --------------------------------------------------------------------------
def chol_upper_jitter(K: np.ndarray):
    """``make_posdef!`` + ``cholesky!(Symmetric(., :U))`` -> (U, info, K_used).

    info = 0: factorised first try; k in 1..10: succeeded after k cumulative
    additions of 1e-6*tr(K)/n to the stored diagonal; -1: still not PD.
    LAPACK dpotrf('U') failure rule (pivot <= 0 or NaN) decides."""
    n = K.shape[0]
    K = np.array(K, dtype=np.float64, order="F", copy=True)
    for attempt in range(MAX_JITTER + 1):
        U, info = lapack.dpotrf(K, lower=0, clean=1, overwrite_a=0)
        if info == 0:
            return U, attempt, K
        if attempt == MAX_JITTER:
            break
        K[np.diag_indices(n)] += 1e-6 * np.trace(K) / n
    return None, -1, K


# --------------------------------------------------------------------------
# log marginal likelihood and gradient (GPE.jl update_mll!, update_mll_and_dmll!)
# --------------------------------------------------------------------------
def eval_mll(X, ymm, theta, kind="se", with_grad=True, return_state=False, diag_offset=0.0):
    """One objective evaluation of one GP.

    X: (n, d); ymm = y - m(X): (n,); theta: (d+2,).
    Returns dict with mll, grad (d+2 or None), info, and optionally the state
    (K, U, alpha, Kinv) for the parity taps.  Non-finite theta -> info -2;
    not PD after 10 jitters -> info -1; in both cases mll = -inf, grad = nan."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    ymm = np.asarray(ymm, dtype=np.float64)
    n, d = X.shape
    theta = np.asarray(theta, dtype=np.float64)
    out = {"mll": -math.inf, "grad": None, "info": 0}
    if with_grad:
        out["grad"] = np.full(d + 2, np.nan)
    if not np.all(np.isfinite(theta)):
        out["info"] = -2
        return out
    log_noise, w, sf2 = unpack_theta(theta, d)
    with np.errstate(over="ignore", invalid="ignore"):
        K0 = assemble_K(X, theta, kind)
    if not np.all(np.isfinite(K0)):
        out["info"] = -2
        return out
    if diag_offset != 0.0:  # gprb_batch_set_diag_offset: a fixed nugget on the stored diagonal (tests of k >= 2 jitters)
        K0[np.diag_indices_from(K0)] += diag_offset
    U, info, K = chol_upper_jitter(K0)
    out["info"] = info
    if U is None:
        return out
    # alpha = cK \ (y - mu)  (dpotrs)
    alpha, _ = lapack.dpotrs(U, ymm, lower=0)
    logdet = 2.0 * float(np.sum(np.log(np.diag(U))))
    mll = -0.5 * (float(ymm @ alpha) + logdet + n * LOG2PI)
    out["mll"] = mll
    state = {"K": K, "U": U, "alpha": alpha, "logdet": logdet}
    if with_grad:
        # get_aainvcKI!: Q = alpha alpha' - K^-1 via dpotrs on -I then ger
        Kinv, _ = lapack.dpotrs(U, np.eye(n), lower=0)
        Q = np.outer(alpha, alpha) - Kinv
        grad = np.empty(d + 2)
        grad[0] = math.exp(2.0 * log_noise) * float(np.trace(Q))  # dmll_noise (no eps)
        r2 = _wsqdist(X, X, w)
        Kf = _kfun(r2, sf2, kind)
        G = Q * _dk_dll_factor(r2, sf2, kind)  # Q o g(r)
        for p in range(d):
            diff = X[:, p][:, None] - X[:, p][None, :]
            grad[1 + p] = 0.5 * w[p] * float(np.sum(G * (diff * diff)))
        grad[d + 1] = float(np.sum(Q * Kf))  # dK/dlsigma = 2 K_f ; 1/2 tr(Q 2K_f)
        out["grad"] = grad
        state["Kinv"] = Kinv
    if return_state:
        out["state"] = state
    return out


def predict(X, theta, state, Xstar, mstar=None, kind="se", want_var=True):
    """``predict_y`` (GP.jl predict_f -> predictMVN; PDMats whiten!).

    mu* = m(x*) + k*' alpha ; var_f = max(k(x*,x*) - ||U^-T k*||^2, 0) ;
    returns (mu, var_f + exp(2 logNoise))."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    Xstar = np.ascontiguousarray(Xstar, dtype=np.float64)
    d = X.shape[1]
    log_noise, _, sf2 = unpack_theta(theta, d)
    Kc = cov_f(X, theta, kind, Xb=Xstar)  # n x m
    mu = Kc.T @ state["alpha"]
    if mstar is not None:
        mu = mu + np.asarray(mstar, dtype=np.float64)
    if not want_var:
        return mu, None
    Lck, _ = lapack.dtrtrs(state["U"], Kc, lower=0, trans=1)  # U' Lck = Kc
    var_f = np.maximum(sf2 - np.sum(Lck * Lck, axis=0), 0.0)
    return mu, var_f + math.exp(2.0 * log_noise)


# --------------------------------------------------------------------------
# extended-precision arbiter (small n only)
# --------------------------------------------------------------------------
def longdouble_eval(X, ymm, theta, kind="se"):
    """Same objective and gradient in numpy.longdouble (x87 80-bit here) with a
    hand-written Cholesky; used to arbitrate when two correct fp64
    factorizations disagree at the cond*eps level.  O(n^3) python-level numpy,
    keep n <= ~300."""
    ld = np.longdouble
    X = np.asarray(X, dtype=ld)
    ymm = np.asarray(ymm, dtype=ld)
    n, d = X.shape
    th = np.asarray(theta, dtype=ld)
    w = np.exp(-2 * th[1 : d + 1])
    sf2 = np.exp(2 * th[d + 1])
    sn2 = np.exp(2 * th[0])
    r2 = np.zeros((n, n), dtype=ld)
    D = []
    for p in range(d):
        diff = X[:, p][:, None] - X[:, p][None, :]
        D.append(diff * diff)
        r2 += w[p] * D[-1]
    r = np.sqrt(r2)
    s3, s5 = np.sqrt(ld(3)), np.sqrt(ld(5))
    if kind == "se":
        Kf = sf2 * np.exp(-r2 / 2)
        g = Kf
    elif kind == "mat12":
        Kf = sf2 * np.exp(-r)
        with np.errstate(divide="ignore", invalid="ignore"):
            g = np.where(r > 0, Kf / r, ld(0))
    elif kind == "mat32":
        Kf = sf2 * (1 + s3 * r) * np.exp(-s3 * r)
        g = 3 * sf2 * np.exp(-s3 * r)
    elif kind == "mat52":
        Kf = sf2 * (1 + s5 * r + 5 * r2 / 3) * np.exp(-s5 * r)
        g = ld(5) / 3 * sf2 * (1 + s5 * r) * np.exp(-s5 * r)
    else:
        raise ValueError(kind)
    K = Kf + (sn2 + ld(EPS)) * np.eye(n, dtype=ld)
    L = np.zeros((n, n), dtype=ld)
    A = K.copy()
    for j in range(n):
        piv = A[j, j] - L[j, :j] @ L[j, :j]
        if not piv > 0:
            raise np.linalg.LinAlgError("not PD in longdouble")
        L[j, j] = np.sqrt(piv)
        if j + 1 < n:
            L[j + 1 :, j] = (A[j + 1 :, j] - L[j + 1 :, :j] @ L[j, :j]) / L[j, j]
    # forward/back substitution against [y, I]
    B = np.concatenate([ymm[:, None], np.eye(n, dtype=ld)], axis=1)
    Z = np.zeros_like(B)
    for i in range(n):
        Z[i] = (B[i] - L[i, :i] @ Z[:i]) / L[i, i]
    Wm = np.zeros_like(B)
    for i in range(n - 1, -1, -1):
        Wm[i] = (Z[i] - L[i + 1 :, i] @ Wm[i + 1 :]) / L[i, i]
    alpha = Wm[:, 0]
    Kinv = Wm[:, 1:]
    logdet = 2 * np.sum(np.log(np.diag(L)))
    mll = -(ymm @ alpha + logdet + n * np.log(2 * ld(np.pi))) / 2
    Q = np.outer(alpha, alpha) - Kinv
    grad = np.zeros(d + 2, dtype=ld)
    grad[0] = sn2 * np.trace(Q)
    G = Q * g
    for p in range(d):
        grad[1 + p] = w[p] * np.sum(G * D[p]) / 2
    grad[d + 1] = np.sum(Q * Kf)
    return {"mll": mll, "grad": grad, "alpha": alpha, "K": K, "Kinv": Kinv, "L": L}


def cond_estimate(K: np.ndarray) -> float:
    ev = np.linalg.eigvalsh(K)
    return float(ev[-1] / max(ev[0], 1e-300))
