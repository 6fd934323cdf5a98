/* libgprb200.so - C ABI of the B200-native GP-regression hot path of GPR.jl.
 *
 * Drop-in boundary: these entry points are what a Julia `ccall`
 * shim bound in place of the four GaussianProcesses.jl calls the reference makes
 * would hit.  Citations are relative to /root/reference.
 *
 *   SEArd(log.(l), log(sf)); GP(X, y, mean, kernel)        examples/maximal_coordinates/CPnoise.jl:38-40
 *       -> gprb_dataset_create (X, shared by the G GPs of a trial)
 *          gprb_batch_create   (y - m(X) per GP; m(.) stays on the host: src/mDynamics.jl:41-55)
 *   update_mll! / update_mll_and_dmll!  (objective inside optimize!, CPnoise.jl:41)
 *       -> gprb_eval           (value-only when grad == NULL)
 *   GaussianProcesses.optimize!(gp, LBFGS(linesearch=BackTracking(order=2)), Optim.Options(...))   CPnoise.jl:41
 *       -> gprb_optimize       (batched lock-step restatement, per-GP masks)
 *   predict_y(gp, Xstar)                                   examples/utils/predictdynamics.jl:13
 *       -> gprb_predict
 *   gp.cK / gp.alpha inspection (parity taps)              -> gprb_get_K / _chol / _alpha / _Kinv
 *
 * Conventions
 *   - All matrices are column-major doubles, exactly Julia's Array{Float64} memory.
 *     X is d x n (one CState sample per column, src/CState.jl:20), Xstar is d x m.
 *   - theta per GP has P = d + 2 entries in GaussianProcesses get_params order:
 *         [logNoise, ll_1 .. ll_d, lsigma]      (log std-dev noise, log length-scales, log signal std)
 *   - Every function returns an int status: 0 = OK, < 0 = error (message via gprb_last_error()).
 *     No exceptions cross the boundary.  Numerical status is per GP in info[b]:
 *         0        factorised first try
 *         1..10    succeeded after that many cumulative jitter additions of 1e-6*tr(K)/n (make_posdef!)
 *         -1       not positive definite after 10 jitters  => mll = -Inf (shim maps to objective +Inf)
 *         -2       non-finite theta / kernel matrix        => mll = -Inf
 *   - Caller owns every host buffer; handles own device memory.  Calls are synchronous on return.
 *   - One gprb_ctx per process and GPU (one process per GPU; ranks are launched by torchrun / Distributed.jl).
 *   - There is NO CPU fallback: gprb_init fails when no sm_100 device is present.
 */
#ifndef GPRB200_H
#define GPRB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gprb_ctx gprb_ctx;
typedef struct gprb_dataset gprb_dataset;
typedef struct gprb_batch gprb_batch;

/* kernel_kind: SEArd is the only family the reference uses (18 call sites);
 * Mat12Ard/Mat32Ard/Mat52Ard are the north_star extension (same theta layout). */
enum { GPRB_KERNEL_SE_ARD = 0, GPRB_KERNEL_MAT12_ARD = 1, GPRB_KERNEL_MAT32_ARD = 2, GPRB_KERNEL_MAT52_ARD = 3 };

enum {
  GPRB_OK = 0,
  GPRB_ERR_ARG = -1,      /* API misuse (null pointer, bad size, mismatched handles) */
  GPRB_ERR_CUDA = -2,     /* CUDA runtime error */
  GPRB_ERR_NODEVICE = -3, /* no sm_100 GPU visible: there is no CPU fallback */
  GPRB_ERR_NOMEM = -4     /* device allocation failed */
};

int gprb_version(void);               /* major*10000 + minor*100 + patch */
const char* gprb_last_error(void);    /* thread-local, never NULL */

/* ---- context ------------------------------------------------------------------------ */
int gprb_init(gprb_ctx** ctx, int device);
int gprb_destroy(gprb_ctx* ctx);
/* Device facts for roofline reporting: out[0]=SM count, out[1]=sm clock kHz, out[2]=L2 bytes, out[3]=free HBM bytes */
int gprb_device_info(gprb_ctx* ctx, int64_t out[4]);

/* ---- dataset: X of one trial, uploaded once (CPnoise.jl:26 `reduce(hcat, ...)`) ------ */
int gprb_dataset_create(gprb_ctx* ctx, int64_t n, int32_t d, const double* X, int64_t ldx, gprb_dataset** out);
int gprb_dataset_update(gprb_dataset* ds, const double* X, int64_t ldx); /* same n, d; new samples */
/* Re-upload `count` datasets in one call: all copies are queued back to back (asynchronous when the host matrices are
 * page-locked) and the call returns after one synchronisation.  X[i] is d x n with leading dimension ldx. */
int gprb_datasets_update(gprb_ctx* ctx, int32_t count, gprb_dataset* const* ds, const double* const* X, int64_t ldx);
int gprb_dataset_destroy(gprb_dataset* ds);

/* ---- batch: B independent GPs, GP b uses dataset ds[b] and targets ymm[:, b] ---------- */
/* ymm = y - m(X), n x B column-major.  All datasets must share n and d. */
int gprb_batch_create(gprb_ctx* ctx, int32_t B, gprb_dataset* const* ds, const double* ymm, int32_t kernel_kind,
                      gprb_batch** out);
int gprb_batch_set_targets(gprb_batch* batch, const double* ymm); /* n x B */
int gprb_batch_destroy(gprb_batch* batch);

/* One objective evaluation per active GP (rows a5-a9 of SURVEY.md section 8a).
 *   theta  P x B    in
 *   active B        in, NULL = all; inactive GPs keep their previous state and outputs are untouched
 *   mll    B        out  log marginal likelihood
 *   grad   P x B    out  d mll / d theta, NULL => value-only (assembly + Cholesky + solve)
 *   info   B        out  see above
 * After return, the factor, alpha and theta of every evaluated GP stay resident for gprb_predict. */
int gprb_eval(gprb_batch* batch, const double* theta, const uint8_t* active, double* mll, double* grad, int32_t* info);

/* Mixed pass: mode[b] = 0 skip, 1 value only, 2 value + gradient.  One pipeline pass evaluates the value-only and the
 * value+gradient GPs together (factorisation for all, inverse + gradient for the mode-2 subset); this is what the
 * batched optimiser issues once per round (line-search trials of some GPs next to the gradient evaluations of the
 * GPs that just accepted a step).  grad may be NULL when no GP asks for mode 2; grad rows of other GPs are untouched. */
int gprb_eval_mixed(gprb_batch* batch, const double* theta, const uint8_t* mode, double* mll, double* grad, int32_t* info);

/* Same evaluation for ALL GPs with theta / outputs already in device memory: no payload crosses PCIe (only the
 * 4-byte per-GP status the make_posdef! retry loop needs).  Ordered after prior work on `stream` (a cudaStream_t);
 * complete on return.  grad_dev / info_dev may be NULL. */
int gprb_eval_device(gprb_batch* batch, const double* theta_dev, double* mll_dev, double* grad_dev, int32_t* info_dev,
                     void* stream);

/* ---- batched L-BFGS + BackTracking(order=2) (Optim 1.4.1 semantics) -- */
typedef struct gprb_lbfgs_opts {
  int32_t m;             /* history, default 10 */
  int32_t iterations;    /* default 1000 */
  int32_t max_evals;     /* 0 = unlimited; deterministic replacement for time_limit */
  int32_t ls_iterations; /* default 1000 */
  double g_abstol;       /* default 1e-8 on ||g||_inf */
  double time_limit;     /* seconds of wall clock for the whole batch, <= 0 = none (reference: 10 s per GP) */
  double c_1, rho_hi, rho_lo; /* 1e-4, 0.5, 0.1 */
} gprb_lbfgs_opts;

typedef struct gprb_opt_result {
  double mll;          /* final log marginal likelihood at theta_inout */
  double g_norm;       /* ||grad||_inf there */
  int32_t iterations;
  int32_t f_calls;     /* value-only evaluations */
  int32_t fg_calls;    /* value+gradient evaluations */
  int32_t converged;   /* 1 = g_abstol / stall criterion met */
  int32_t ls_failed;   /* 1 = LineSearchException equivalent */
  int32_t info;        /* info of the final evaluation */
} gprb_opt_result;

void gprb_lbfgs_default_opts(gprb_lbfgs_opts* o);
/* theta_inout P x B: start point in, minimiser of -mll out.  Final state is left evaluated at the minimiser
 * (mirrors optimize! writing the minimiser back and calling update_target!). */
int gprb_optimize(gprb_batch* batch, double* theta_inout, const gprb_lbfgs_opts* opts, gprb_opt_result* results);

/* Host-only self test of the batched optimiser state machine (no GPU needed): B copies of the P-dimensional
 * Rosenbrock function, +Inf outside |x_i| <= bound.  results[b].mll carries the final objective value.
 * Test hook for tests/test_lbfgs_cpu.py; not part of the reference-facing surface. */
int gprb_lbfgs_selftest(int32_t B, int32_t P, double* theta_inout, const gprb_lbfgs_opts* opts, double bound,
                        gprb_opt_result* results);

/* ---- prediction (predict_y) ------------------------------------------------------------ */
/* Xstar: d x m column-major.  xstar_stride = 0: the same d x m block is used by every GP; otherwise GP b reads
 * its own block at Xstar + b*xstar_stride doubles (xstar_stride >= d*m).
 *   mstar m x B  prior mean m(x*) evaluated on the host, NULL = zero mean
 *   mu    m x B  out
 *   var   m x B  out, NULL => mean only (the reference discards the variance: predictdynamics.jl:13) */
int gprb_predict(gprb_batch* batch, int64_t m, const double* Xstar, int64_t xstar_stride, const double* mstar,
                 double* mu, double* var);

/* ---- parity taps (state after the last gprb_eval of GP b) ------------------------------ */
int gprb_get_K(gprb_batch* batch, int32_t b, double* out /* n x n, symmetric, noise + jitter included */);
int gprb_get_chol(gprb_batch* batch, int32_t b, double* out /* n x n upper U with K = U'U, zeros below */);
int gprb_get_alpha(gprb_batch* batch, int32_t b, double* out /* n */);
int gprb_get_Kinv(gprb_batch* batch, int32_t b, double* out /* n x n symmetric; needs a value+grad eval */);

/* ---- timing hooks for bench.py (CUDA events on the library's own streams) -------------- */
/* Per-stage device time of the most recent gprb_eval / gprb_eval_device, milliseconds:
 * out[0]=assembly out[1]=cholesky out[2]=solve+mll out[3]=inverse out[4]=gradient out[5]=total
 * out[6]=sum over the DMMA tile-GEMM launches alone, out[7]=number of those launches.
 * Valid only when profiling was enabled with gprb_set_profiling(batch, 1) (single stream, stages serialised). */
int gprb_set_profiling(gprb_batch* batch, int32_t on);
int gprb_last_stage_ms(gprb_batch* batch, double out[8]);
/* Device time of every tile-GEMM launch of the last profiled evaluation, in launch order (J x [CHOL_DIAG, CHOL_COL],
 * then TRTRI_ROW 1..J-1, then LAUUM).  Writes at most `cap` values, returns the number of launches (>= 0). */
int gprb_last_gemm_launch_ms(gprb_batch* batch, double* out, int32_t cap);
/* Number of kernel launches issued by the library since the context was created. */
int64_t gprb_launch_count(gprb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* GPRB200_H */
