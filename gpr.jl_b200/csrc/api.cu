// C ABI of libgprb200 (include/gprb200.h): handle management and the host-side orchestration of the
// evaluation pipeline  assembly -> blocked Cholesky -> solves/mll -> inverse -> fused gradient.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <limits>
#include <new>

#include "common.cuh"
#include "kernels.h"

namespace gprb {

static thread_local std::string g_err = "";
void set_error(const std::string& msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" + std::to_string(line) + ")";
  return e == cudaErrorMemoryAllocation ? GPRB_ERR_NOMEM : GPRB_ERR_CUDA;
}

// fresh evaluation: the diagonal term starts at the GP's fixed offset (gprb_batch_set_diag_offset, default 0)
__global__ void k_reset_jitter(double* jitter, const double* diag_off, const int32_t* list, int count) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) jitter[list[k]] = diag_off[list[k]];
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
int make_tensor_map(CUtensorMap* map, const double* base, uint64_t rows, uint64_t cols, uint64_t gps, uint64_t col_stride_doubles,
                    uint64_t gp_stride_doubles, uint32_t box_rows, uint32_t box_cols) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled is not available from this driver (TMA tensor maps are required)");
      return GPRB_ERR_CUDA;
    }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t gdim[3] = {rows, cols, gps};
  const cuuint64_t gstride[2] = {col_stride_doubles * sizeof(double), gp_stride_doubles * sizeof(double)};
  const cuuint32_t box[3] = {box_rows, box_cols, 1};
  const cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), gdim, gstride, box, estride,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return GPRB_ERR_CUDA;
  }
  return 0;
}

static void set_gemm_maps(GemmArgs& ga, const gprb_batch* b) {
  ga.tm_L132 = b->tm_L132; ga.tm_L68 = b->tm_L68; ga.tm_A68 = b->tm_A68; ga.tm_DT132 = b->tm_DT132; ga.tm_DT68 = b->tm_DT68; ga.tm_D132 = b->tm_D132;
}

template <typename T>
static int dev_alloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
  return 0;
}

static void free_batch(gprb_batch* b) {
  if (!b) return;
  cudaFree(b->Xptr); cudaFree(b->Xtptr); cudaFree(b->ymm); cudaFree(b->theta); cudaFree(b->A); cudaFree(b->Lm);
  cudaFree(b->Dinv); cudaFree(b->DinvT); cudaFree(b->KinvD); cudaFree(b->alpha); cudaFree(b->zbuf); cudaFree(b->jitter); cudaFree(b->diag_off);
  cudaFree(b->logdet_part); cudaFree(b->fail); cudaFree(b->mll); cudaFree(b->grad);
  cudaFree(b->grad_part); cudaFree(b->list);
  for (gprb_predict_slot& sl : b->ps) {
    cudaFree(sl.pX); cudaFree(sl.pms); cudaFree(sl.pmu); cudaFree(sl.pvar); cudaFree(sl.pT); cudaFree(sl.pmupart);
    cudaFree(sl.pq); cudaFree(sl.pcount); cudaFree(sl.mask);
    if (sl.h_in) cudaFreeHost(sl.h_in);
    if (sl.h_out) cudaFreeHost(sl.h_out);
    if (sl.h_mask) cudaFreeHost(sl.h_mask);
    if (sl.t0) cudaEventDestroy(sl.t0);
    if (sl.t1) cudaEventDestroy(sl.t1);
  }
  if (b->list_host) cudaFreeHost(b->list_host);
  if (b->fail_host) cudaFreeHost(b->fail_host);
  if (b->stage_host) cudaFreeHost(b->stage_host);
  for (int s = 0; s < MAX_STREAMS; ++s) {
    if (b->stream[s]) cudaStreamDestroy(b->stream[s]);
    if (b->join[s]) cudaEventDestroy(b->join[s]);
  }
  for (int s = 0; s < 8; ++s)
    if (b->ev[s]) cudaEventDestroy(b->ev[s]);
  for (cudaEvent_t e : b->gemm_ev) cudaEventDestroy(e);
  delete b;
}

// Enqueue one evaluation of the GPs in list[off .. off+count) on `st`: inverse + fused gradient for the first `ngrad`
// entries of the segment (the host orders value+gradient GPs first); assembly, Cholesky and solve for all entries but
// the first `nreuse` (<= ngrad), whose factor, alpha and mll of the previous evaluation at the same theta are still
// resident (the optimiser asks for the gradient at the point its line search just accepted).
// `ops` != nullptr: the launches are not issued but appended (as closures) to `ops`, so that run_pipeline can issue the
// launches of all stream groups interleaved - every group then starts at once instead of one enqueue time (0.3 ms of host
// work per group) after the previous one, which matters for short passes (50 GPs per GPU in the 8-GPU strong split).
using LaunchOps = std::vector<std::function<int()>>;
static int enqueue_pipeline(gprb_batch* b, int off, int count, int ngrad, int nreuse, bool right_looking, cudaStream_t st,
                            bool prof, LaunchOps* ops = nullptr) {
  if (count <= 0) return 0;
  const bool with_grad = ngrad > 0;
  const int32_t* glist = b->list + off;     // inverse + gradient: glist[0 .. ngrad)
  const int32_t* list = glist + nreuse;     // assembly, factorisation, solves: list[0 .. count)
  count -= nreuse;
  const int J = b->J;
  const int64_t ms = b->npad * b->npad, dstride = (int64_t)J * NB * NB;
  int rc;
  int64_t& launches = b->ctx->launches;
  int gcount = count;  // GPs per tile-GEMM launch (count for the factorisation, ngrad for the inverse)
  auto run = [&](std::function<int()> f) -> int {
    if (ops) { ops->push_back(std::move(f)); return 0; }
    return f();
  };
  auto gemm = [&](const GemmArgs& a, int ntiles) -> int {
    const int count = gcount;
    if (!prof) return run([a, ntiles, count, st] { return launch_tile_gemm(a, ntiles, count, st); });
    while ((int)b->gemm_ev.size() < b->gemm_ev_used + 2) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaEventCreate", __FILE__, __LINE__);
      b->gemm_ev.push_back(e);
    }
    cudaEventRecord(b->gemm_ev[b->gemm_ev_used], st);
#ifdef GPRB_TIMELINE
    // debug build only: per-tile phase timestamps of this launch, averaged and printed to stderr
    static unsigned long long* tl_dev = nullptr;
    const size_t tl_n = (size_t)count * ntiles * 2 * 8;  // two half-tile CTAs per logical tile (FWD_ROW: one; the rest stays zero)
    if (!tl_dev) cudaMalloc(&tl_dev, sizeof(unsigned long long) * 8 * 65536 * 16);
    cudaMemsetAsync(tl_dev, 0, sizeof(unsigned long long) * tl_n, st);
    GemmArgs a2 = a; a2.tl = tl_dev;
    int r = launch_tile_gemm(a2, ntiles, count, st);
    cudaEventRecord(b->gemm_ev[b->gemm_ev_used + 1], st);
    b->gemm_ev_used += 2;
    {
      std::vector<unsigned long long> h(tl_n);
      cudaStreamSynchronize(st);
      cudaMemcpy(h.data(), tl_dev, sizeof(unsigned long long) * tl_n, cudaMemcpyDeviceToHost);
      double ph[6] = {0, 0, 0, 0, 0, 0}, tot = 0, waitc = 0; size_t cnt = 0;
      for (size_t k = 0; k < tl_n; k += 8) {
        const unsigned long long* t = &h[k];
        if (!t[0] || !t[6]) continue;
        unsigned long long prev = t[0];
        for (int q = 1; q <= 6; ++q) { unsigned long long cur = t[q] ? t[q] : prev; ph[q - 1] += (double)(cur - prev); prev = cur; }
        tot += (double)(t[6] - t[0]); waitc += (double)t[7]; ++cnt;
      }
      if (cnt) fprintf(stderr, "TL mode %d step %2d tiles %6zu  fill %6.2f main %7.2f cin %6.2f park %6.2f post %6.2f store %6.2f  total %7.2f us  (operand wait in main, after the first chunk %6.2f us)\n",
                       a.mode, a.step, cnt, ph[0] / cnt / 1e3, ph[1] / cnt / 1e3, ph[2] / cnt / 1e3, ph[3] / cnt / 1e3, ph[4] / cnt / 1e3, ph[5] / cnt / 1e3, tot / cnt / 1e3, waitc / cnt / 1965.0);
    }
    return r;
#else
    int r = launch_tile_gemm(a, ntiles, count, st);
    cudaEventRecord(b->gemm_ev[b->gemm_ev_used + 1], st);
    b->gemm_ev_used += 2;
    return r;
#endif
  };
  if (prof) { b->gemm_ev_used = 0; cudaEventRecord(b->ev[0], st); }
  const int nv = (int)((b->n + KT - 1) / KT * KT);
  GemmArgs ga{b->Lm, b->DinvT, b->Dinv, b->A, b->Lm, b->KinvD, list, ms, dstride, (int)b->npad, J, 0, GEMM_CHOL_DIAG, nv};
  ga.fail = b->fail;
  set_gemm_maps(ga, b);
  { static const int pf = getenv("GPRB200_PF_CIN") ? atoi(getenv("GPRB200_PF_CIN")) : 1; ga.pf_cin = (!right_looking && pf) ? 1 : 0; }
  if (count > 0) {
  // Small passes are latency bound (one dependent chain of 3 J launches): they use the right-looking factorisation,
  // whose launches are short and wide, instead of the left-looking one, whose k-loops grow with the column index.
  AssembleArgs aa{b->Xtptr, b->theta, b->jitter, b->A, nullptr, b->fail, list, ms, (int)b->n, (int)b->npad, b->d, J, b->kind};
  if (right_looking) aa.A2 = b->Lm;
  if ((rc = run([aa, count, st] { return launch_assemble(aa, count, st); }))) return rc;
  ++launches;
  if (prof) cudaEventRecord(b->ev[1], st);
  DiagArgs da{nullptr, b->Lm, b->Dinv, b->DinvT, b->logdet_part, b->fail, list, ms, dstride, (int)b->npad, J, 0, nv};
  if (right_looking) ga.Cin = b->Lm;  // S lives (and is updated in place) in the lower tiles of Lm
  for (int j = 0; j < J; ++j) {
    ga.step = j;
    if (!right_looking && j > 0) {  // S(0,0) = K(0,0): k_diag_factor reads block column 0 straight from A
      ga.mode = GEMM_CHOL_DIAG;
      if ((rc = gemm(ga, 1))) return rc;
      ++launches;
    }
    da.step = j;
    da.Src = (!right_looking && j == 0) ? b->A : nullptr;
    if ((rc = run([da, count, st] { return launch_diag_factor(da, count, st); }))) return rc;
    ++launches;
    if (j + 1 < J) {
      ga.mode = right_looking ? GEMM_CHOL_PANEL : GEMM_CHOL_COL;
      if ((rc = gemm(ga, J - 1 - j))) return rc;
      ++launches;
      if (right_looking) {
        ga.mode = GEMM_CHOL_TRAIL;
        if ((rc = gemm(ga, (J - 1 - j) * (J - j) / 2))) return rc;
        ++launches;
      }
    }
  }
  if (prof) cudaEventRecord(b->ev[2], st);
  SolveArgs sa{b->Lm, b->Dinv, b->ymm, b->logdet_part, b->fail, b->zbuf, b->alpha, b->mll, list, ms, dstride,
               (int)b->n, (int)b->npad, J, nv};
  sa.cluster_below = b->solve_cluster_below;
  if ((rc = run([sa, count, st] { return launch_solve(sa, count, st); }))) return rc;
  ++launches;
  }
  if (prof) cudaEventRecord(b->ev[3], st);
  if (with_grad) {
    ga.Cin = nullptr;
    ga.list = glist;
    gcount = ngrad;
    for (int i = 1; i < J; ++i) {
      ga.step = i; ga.mode = GEMM_TRTRI_ROW; ga.Cout = b->Lm;
      if ((rc = gemm(ga, i))) return rc;
      ++launches;
    }
    ga.mode = GEMM_LAUUM; ga.step = 0; ga.Cout = b->A;
    if ((rc = gemm(ga, J * (J + 1) / 2))) return rc;
    ++launches;
    if (prof) cudaEventRecord(b->ev[4], st);
    GradArgs gr{b->Xtptr, b->theta, b->A, b->KinvD, b->alpha, b->grad_part, b->grad, b->fail, glist, ms, dstride,
                (int)b->n, (int)b->npad, b->d, J, b->kind};
    if ((rc = run([gr, ngrad, st] { return launch_grad(gr, ngrad, st); }))) return rc;
    launches += 2;
    if (prof) cudaEventRecord(b->ev[5], st);
  }
  return 0;
}

// One stream group of an evaluation pass: list[off .. off+count), the first ngrad of them with gradient, the first
// nreuse of those on the resident factorisation.
struct Group { int off, count, ngrad, nreuse; bool rl; };

// Order the active GPs of a pass into stream groups inside list_host: the value+gradient GPs and the value-only GPs
// are each dealt evenly over the groups (equal work per stream), gradient GPs first inside every group.
static std::vector<Group> build_groups(gprb_batch* b, const std::vector<int32_t>& reuse_gps, const std::vector<int32_t>& grad_gps,
                                       const std::vector<int32_t>& val_gps) {
  const int total = (int)(reuse_gps.size() + grad_gps.size() + val_gps.size());
  const int S = (b->profiling || total < 8 || total <= b->rl_max || b->nstreams == 1) ? 1 : b->nstreams;
  std::vector<Group> groups;
  int off = 0;
  // Unequal group sizes (weights 1 + skew, ..., 1 - skew): equal groups start together and run the same launches, so
  // they reach their low-occupancy phases (diagonal tiles, potf2, the first rows of the triangular inverse) at the same
  // time; groups of different size drift apart, and the thin phases of one fall under the wide phases of another.
  std::vector<double> cum(S + 1, 0.0);
  for (int s = 0; s < S; ++s) cum[s + 1] = cum[s] + (S > 1 ? 1.0 + b->group_skew * (double)(S - 1 - 2 * s) / (double)(S - 1) : 1.0);
  auto cut = [&](size_t m, int s) { return (int)llround((double)m * cum[s] / cum[S]); };
  for (int s = 0; s < S; ++s) {
    const int r0 = cut(reuse_gps.size(), s), r1 = cut(reuse_gps.size(), s + 1);
    const int g0 = cut(grad_gps.size(), s), g1 = cut(grad_gps.size(), s + 1);
    const int v0 = cut(val_gps.size(), s), v1 = cut(val_gps.size(), s + 1);
    Group g{off, (r1 - r0) + (g1 - g0) + (v1 - v0), (r1 - r0) + (g1 - g0), r1 - r0, !b->profiling && total <= b->rl_max};
    for (int k = r0; k < r1; ++k) b->list_host[off++] = reuse_gps[k];
    for (int k = g0; k < g1; ++k) b->list_host[off++] = grad_gps[k];
    for (int k = v0; k < v1; ++k) b->list_host[off++] = val_gps[k];
    if (g.count > 0) groups.push_back(g);
  }
  return groups;
}

// Run one pass: every group on its own stream so the serial diagonal-block kernels of one group overlap the DMMA
// tiles of another.  Joins everything on stream[0].
static int run_pipeline(gprb_batch* b, const std::vector<Group>& groups) {
  int rc;
  if (groups.empty()) return 0;
  if (b->profiling) {
    const Group& g = groups[0];
    if ((rc = enqueue_pipeline(b, g.off, g.count, g.ngrad, g.nreuse, g.rl, b->stream[0], true))) return rc;
    GPRB_CUDA(cudaStreamSynchronize(b->stream[0]));
    float ms = 0.f;
    const int last = g.ngrad > 0 ? 5 : 3;
    for (int s = 0; s < 5; ++s) b->stage_ms[s] = 0.0;
    for (int s = 0; s < last; ++s) {
      GPRB_CUDA(cudaEventElapsedTime(&ms, b->ev[s], b->ev[s + 1]));
      b->stage_ms[s] = ms;
    }
    GPRB_CUDA(cudaEventElapsedTime(&ms, b->ev[0], b->ev[last]));
    b->stage_ms[5] = ms;
    double gsum = 0.0;
    b->gemm_ms.clear();
    for (int k = 0; k + 1 < b->gemm_ev_used; k += 2) {
      GPRB_CUDA(cudaEventElapsedTime(&ms, b->gemm_ev[k], b->gemm_ev[k + 1]));
      gsum += ms;
      b->gemm_ms.push_back(ms);
    }
    b->stage_ms[6] = gsum;
    b->stage_ms[7] = b->gemm_ev_used / 2;
    return 0;
  }
  GPRB_CUDA(cudaEventRecord(b->join[0], b->stream[0]));
  std::vector<LaunchOps> ops(groups.size());
  size_t longest = 0;
  for (size_t s = 0; s < groups.size(); ++s) {
    const Group& g = groups[s];
    if (s > 0) GPRB_CUDA(cudaStreamWaitEvent(b->stream[s], b->join[0], 0));
    if ((rc = enqueue_pipeline(b, g.off, g.count, g.ngrad, g.nreuse, g.rl, b->stream[s], false, groups.size() > 1 ? &ops[s] : nullptr)))
      return rc;
    longest = std::max(longest, ops[s].size());
  }
  for (size_t k = 0; k < longest; ++k)  // launch k of every group, group by group: all groups advance together
    for (size_t s = 0; s < groups.size(); ++s)
      if (k < ops[s].size() && (rc = ops[s][k]())) return rc;
  for (size_t s = 1; s < groups.size(); ++s) {
    GPRB_CUDA(cudaEventRecord(b->join[s], b->stream[s]));
    GPRB_CUDA(cudaStreamWaitEvent(b->stream[0], b->join[s], 0));
  }
  return 0;
}

// ONE pipeline pass over the GPs with mode != 0 (theta already on the device).  retry[gp] != 0: the GP's previous
// pass failed its factorisation - its cumulative jitter grows by 1e-6 tr(K)/n (make_posdef!) instead of being reset.
// Afterwards: info[gp] final for the GPs that are done, pending[gp] = 1 for those that need another retry pass.
//
// State reuse (theta_host != nullptr): a GP asked to evaluate, first try, at exactly the theta of its last successful
// evaluation (same dataset upload) keeps what is resident - a value-only request is answered from the resident mll, a
// gradient request only runs the inverse and the fused gradient on the resident factor (or nothing, when that gradient
// is resident too).  The results are bit-identical to a fresh evaluation: every stage is deterministic, and the jitter
// the earlier evaluation settled on is the one make_posdef! would arrive at again.  This is the optimiser's pattern:
// the reference evaluates the point a line search accepts twice (value, then value + gradient), and once more after
// the last iteration.
static int eval_pass(gprb_batch* b, const double* theta_host, const uint8_t* mode, const uint8_t* retry,
                     std::vector<int32_t>& info, std::vector<uint8_t>& pending) {
  int rc;
  const int B = b->B, P = b->P;
  if ((int)b->tries.size() != B) b->tries.assign(B, 0);
  if ((int)b->theta_valid.size() != B) {
    b->theta_valid.assign(B, 0); b->theta_last.assign((size_t)B * P, 0.0); b->info_last.assign(B, 0); b->ds_ver.assign(B, 0);
  }
  pending.assign(B, 0);
  static const bool allow_reuse = [] { const char* e = getenv("GPRB200_REUSE"); return !(e && e[0] == '0'); }();
  std::vector<int32_t> reuse_gps, grad_gps, val_gps, cached;
  int nfresh = 0, nretry = 0;
  int32_t* aux = b->list_host + B;  // fresh GPs from the front, retried GPs from the back
  for (int i = 0; i < B; ++i) {
    if (!mode[i]) continue;
    const bool is_retry = retry && retry[i];
    if (allow_reuse && theta_host && !is_retry && !b->profiling && b->state_ok[i] && b->theta_valid[i] &&
        b->ds_ver[i] == b->ds[i]->version &&
        memcmp(theta_host + (size_t)i * P, b->theta_last.data() + (size_t)i * P, sizeof(double) * P) == 0) {
      if (mode[i] == 1 || b->inv_ok[i]) cached.push_back(i);  // everything asked for is resident
      else reuse_gps.push_back(i);                             // gradient on the resident factor
      continue;
    }
    (mode[i] == 2 ? grad_gps : val_gps).push_back(i);
    if (is_retry) aux[B - 1 - nretry++] = i;
    else { aux[nfresh++] = i; b->tries[i] = 0; }
  }
  for (int gp : cached) info[gp] = b->info_last[gp];
  const int count = (int)(reuse_gps.size() + grad_gps.size() + val_gps.size());
  if (count == 0) return 0;
  const std::vector<Group> groups = build_groups(b, reuse_gps, grad_gps, val_gps);
  GPRB_CUDA(cudaMemcpyAsync(b->list, b->list_host, sizeof(int32_t) * 2 * B, cudaMemcpyHostToDevice, b->stream[0]));
  if (nfresh) {
    k_reset_jitter<<<(nfresh + 127) / 128, 128, 0, b->stream[0]>>>(b->jitter, b->diag_off, b->list + B, nfresh);
    GPRB_CUDA(cudaGetLastError());
  }
  if (nretry && (rc = launch_add_jitter(b->theta, b->jitter, b->list + 2 * B - nretry, b->d, nretry, b->stream[0]))) return rc;
  if ((rc = run_pipeline(b, groups))) return rc;
  GPRB_CUDA(cudaMemcpyAsync(b->fail_host, b->fail, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, b->stream[0]));
  GPRB_CUDA(cudaStreamSynchronize(b->stream[0]));
  if (getenv("GPRB200_LBFGS_TRACE")) {  // where do failed factorisations break down? (first bad pivot / n)
    double sum = 0; int nf = 0;
    for (int gp = 0; gp < B; ++gp)
      if (mode[gp] && b->fail_host[gp] > 0) { sum += (double)b->fail_host[gp] / (double)b->n; ++nf; }
    if (nf) fprintf(stderr, "  pass: %d of %d factorisations failed, mean failing pivot at %.2f n\n", nf, count, sum / nf);
  }
  for (int gp : reuse_gps) { info[gp] = b->info_last[gp]; b->inv_ok[gp] = 1; b->v_ok[gp] = 1; }
  for (int pass = 0; pass < 2; ++pass)
    for (int gp : (pass == 0 ? grad_gps : val_gps)) {
      const int f = b->fail_host[gp];
      b->theta_valid[gp] = 0;
      if (f == 0) info[gp] = b->tries[gp];
      else if (f < 0) info[gp] = -2;
      else if (b->tries[gp] >= MAX_JITTER) info[gp] = -1;
      else { b->tries[gp]++; pending[gp] = 1; b->state_ok[gp] = 0; b->inv_ok[gp] = 0; b->v_ok[gp] = 0; continue; }
      b->state_ok[gp] = info[gp] >= 0;
      b->inv_ok[gp] = pass == 0 && info[gp] >= 0;
      b->v_ok[gp] = b->inv_ok[gp];
      if (theta_host && info[gp] >= 0) {  // remember what is resident now
        memcpy(b->theta_last.data() + (size_t)gp * P, theta_host + (size_t)gp * P, sizeof(double) * P);
        b->theta_valid[gp] = 1;
        b->info_last[gp] = info[gp];
        b->ds_ver[gp] = b->ds[gp]->version;
      }
    }
  return 0;
}

// Evaluation with the make_posdef! retry loop inside (gprb_eval / gprb_eval_mixed / gprb_eval_device): passes are
// repeated for the GPs that still need a jitter until none is pending.
static int evaluate_modes(gprb_batch* b, const double* theta_host, const uint8_t* mode, std::vector<int32_t>& info) {
  info.assign(b->B, 0);
  std::vector<uint8_t> cur(mode, mode + b->B), retry(b->B, 0), pending;
  for (;;) {
    int rc = eval_pass(b, theta_host, cur.data(), retry.data(), info, pending);
    if (rc) return rc;
    bool any = false;
    for (int i = 0; i < b->B; ++i) {
      cur[i] = pending[i] ? cur[i] : 0;
      retry[i] = pending[i];
      any = any || pending[i];
    }
    if (!any) return 0;
  }
}

// theta rows of the listed GPs -> device (pinned staging, one strided copy per contiguous run of GP indices)
static int upload_theta_rows(gprb_batch* b, const double* theta, const std::vector<int32_t>& act) {
  const int P = b->P, count = (int)act.size();
  double* st = b->stage_host;
  for (int k = 0; k < count; ++k) memcpy(st + (size_t)k * P, theta + (size_t)act[k] * P, sizeof(double) * P);
  for (int k = 0; k < count;) {
    int k2 = k + 1;
    while (k2 < count && act[k2] == act[k2 - 1] + 1) ++k2;
    GPRB_CUDA(cudaMemcpyAsync(b->theta + (size_t)act[k] * P, st + (size_t)k * P, sizeof(double) * P * (k2 - k),
                              cudaMemcpyHostToDevice, b->stream[0]));
    k = k2;
  }
  return 0;
}

// mll (+ grad rows of the mode-2 GPs) of the finished GPs -> host arrays
static int download_results(gprb_batch* b, const uint8_t* mode, const std::vector<int32_t>& info, const uint8_t* skip,
                            double* mll, double* grad, int32_t* info_out) {
  const int B = b->B, P = b->P;
  bool any_grad = false;
  for (int i = 0; i < B; ++i) any_grad = any_grad || (mode[i] == 2 && !(skip && skip[i]));
  double* res = b->stage_host;  // [B] mll then [B*P] grad
  GPRB_CUDA(cudaMemcpyAsync(res, b->mll, sizeof(double) * B, cudaMemcpyDeviceToHost, b->stream[0]));
  if (any_grad) GPRB_CUDA(cudaMemcpyAsync(res + B, b->grad, sizeof(double) * B * P, cudaMemcpyDeviceToHost, b->stream[0]));
  GPRB_CUDA(cudaStreamSynchronize(b->stream[0]));
  const double ninf = -std::numeric_limits<double>::infinity(), qnan = std::numeric_limits<double>::quiet_NaN();
  for (int gp = 0; gp < B; ++gp) {
    if (!mode[gp] || (skip && skip[gp])) continue;
    info_out[gp] = info[gp];
    const bool ok = info[gp] >= 0;
    mll[gp] = ok ? res[gp] : ninf;
    if (mode[gp] == 2)
      for (int p = 0; p < P; ++p) grad[(size_t)gp * P + p] = ok ? res[B + (size_t)gp * P + p] : qnan;
  }
  return 0;
}

int eval_pass_host(gprb_batch* b, const double* theta, const uint8_t* mode, const uint8_t* retry, double* mll, double* grad,
                   int32_t* info_out, uint8_t* pending_out) {
  GPRB_REQUIRE(b && theta && mode && mll && grad && info_out && pending_out, "eval_pass_host: NULL argument");
  GPRB_CUDA(cudaSetDevice(b->ctx->device));
  std::vector<int32_t> fresh;
  for (int i = 0; i < b->B; ++i)
    if (mode[i] && !(retry && retry[i])) fresh.push_back(i);
  int rc = upload_theta_rows(b, theta, fresh);
  if (rc) return rc;
  std::vector<int32_t> info(b->B, 0);
  std::vector<uint8_t> pending;
  if ((rc = eval_pass(b, theta, mode, retry, info, pending))) return rc;
  memcpy(pending_out, pending.data(), b->B);
  return download_results(b, mode, info, pending_out, mll, grad, info_out);
}

}  // namespace gprb

using namespace gprb;

extern "C" {

int gprb_version(void) { return 200; }
const char* gprb_last_error(void) { return g_err.c_str(); }

int gprb_init(gprb_ctx** out, int device) {
  GPRB_REQUIRE(out != nullptr, "gprb_init: ctx out-pointer is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("gprb_init: no CUDA device visible - libgprb200 has no CPU fallback");
    return GPRB_ERR_NODEVICE;
  }
  GPRB_REQUIRE(device >= 0 && device < ndev, "gprb_init: device index out of range");
  cudaDeviceProp prop;
  GPRB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error(std::string("gprb_init: device '") + prop.name + "' is not sm_100 (B200); the library is built for sm_100a only");
    return GPRB_ERR_NODEVICE;
  }
  GPRB_CUDA(cudaSetDevice(device));
  // dynamic shared-memory opt-ins are per-device function attributes: set them for THIS device (a second context on
  // another GPU of the same process gets its own)
  int rc;
  if ((rc = configure_tile_gemm()) || (rc = configure_diag_factor())) return rc;
  gprb_ctx* c = new (std::nothrow) gprb_ctx();
  GPRB_REQUIRE(c != nullptr, "gprb_init: out of host memory");
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  c->clock_khz = khz;
  c->l2_bytes = prop.l2CacheSize;
  cudaError_t se = cudaStreamCreateWithFlags(&c->upload, cudaStreamNonBlocking);
  if (se == cudaSuccess) se = cudaStreamCreateWithFlags(&c->upload2, cudaStreamNonBlocking);
  if (se == cudaSuccess) se = cudaEventCreateWithFlags(&c->upload_ev, cudaEventDisableTiming);
  if (se != cudaSuccess) { delete c; return cuda_fail(se, "cudaStreamCreate(upload)", __FILE__, __LINE__); }
  *out = c;
  return GPRB_OK;
}

int gprb_destroy(gprb_ctx* ctx) {
  if (ctx) {
    cudaSetDevice(ctx->device);
    comm_release(ctx);
    if (ctx->upload) cudaStreamDestroy(ctx->upload);
    if (ctx->upload2) cudaStreamDestroy(ctx->upload2);
    if (ctx->upload_ev) cudaEventDestroy(ctx->upload_ev);
  }
  delete ctx;
  return GPRB_OK;
}

int gprb_device_info(gprb_ctx* ctx, int64_t out[4]) {
  GPRB_REQUIRE(ctx && out, "gprb_device_info: NULL argument");
  GPRB_CUDA(cudaSetDevice(ctx->device));
  size_t fr = 0, tot = 0;
  GPRB_CUDA(cudaMemGetInfo(&fr, &tot));
  out[0] = ctx->sm_count; out[1] = ctx->clock_khz; out[2] = ctx->l2_bytes; out[3] = (int64_t)fr;
  return GPRB_OK;
}

int64_t gprb_launch_count(gprb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------------
// Enqueue the host -> device copy of one dataset on `st` (no synchronisation).
static int dataset_copy_async(gprb_dataset* ds, const double* X, int64_t ldx, cudaStream_t st) {
  ds->version++;  // evaluations on the previous contents are no longer reusable
  if (ldx == ds->d)
    GPRB_CUDA(cudaMemcpyAsync(ds->X, X, sizeof(double) * ds->d * ds->n, cudaMemcpyHostToDevice, st));
  else
    GPRB_CUDA(cudaMemcpy2DAsync(ds->X, sizeof(double) * ds->d, X, sizeof(double) * ldx, sizeof(double) * ds->d, ds->n,
                                cudaMemcpyHostToDevice, st));
  return 0;
}

// Enqueue the upload of one dataset (copy + transposed copy Xt) on the context's upload stream (no synchronisation).
static int dataset_upload_async(gprb_dataset* ds, const double* X, int64_t ldx) {
  cudaStream_t st = ds->ctx->upload;
  int rc = dataset_copy_async(ds, X, ldx, st);
  if (rc) return rc;
  if ((rc = launch_transpose_inputs(ds->X, ds->Xt, (int)ds->n, (int)ds->npad, ds->d, st))) return rc;
  ds->ctx->launches++;
  return 0;
}

static int dataset_upload(gprb_dataset* ds, const double* X, int64_t ldx) {
  GPRB_CUDA(cudaSetDevice(ds->ctx->device));
  int rc = dataset_upload_async(ds, X, ldx);
  if (rc) return rc;
  GPRB_CUDA(cudaStreamSynchronize(ds->ctx->upload));
  return 0;
}

int gprb_dataset_create(gprb_ctx* ctx, int64_t n, int32_t d, const double* X, int64_t ldx, gprb_dataset** out) {
  GPRB_REQUIRE(ctx && X && out, "gprb_dataset_create: NULL argument");
  *out = nullptr;
  GPRB_REQUIRE(n >= 1 && n <= (1 << 20), "gprb_dataset_create: n out of range");
  GPRB_REQUIRE(d >= 1 && d <= MAX_D, "gprb_dataset_create: d must be in 1..62");
  GPRB_REQUIRE(ldx >= d, "gprb_dataset_create: ldx < d");
  GPRB_CUDA(cudaSetDevice(ctx->device));
  gprb_dataset* ds = new (std::nothrow) gprb_dataset();
  GPRB_REQUIRE(ds != nullptr, "gprb_dataset_create: out of host memory");
  ds->ctx = ctx; ds->n = n; ds->d = d;
  ds->npad = (n + NB - 1) / NB * NB;
  int rc;
  if ((rc = dev_alloc(&ds->X, (size_t)n * d)) || (rc = dev_alloc(&ds->Xt, (size_t)ds->npad * d)) ||
      (rc = dataset_upload(ds, X, ldx))) {
    cudaFree(ds->X); cudaFree(ds->Xt); delete ds;
    return rc;
  }
  *out = ds;
  return GPRB_OK;
}

int gprb_dataset_update(gprb_dataset* ds, const double* X, int64_t ldx) {
  GPRB_REQUIRE(ds && X, "gprb_dataset_update: NULL argument");
  GPRB_REQUIRE(ldx >= ds->d, "gprb_dataset_update: ldx < d");
  return dataset_upload(ds, X, ldx);
}

// ds[0 .. count) are the members 0 .. count-1 of one slab, in order, and the host matrices form one contiguous block:
// the whole upload is ONE host->device copy and ONE transpose launch (the per-dataset form costs the host ~10 us per
// copy / launch to enqueue - 1 ms per 100 trial datasets, which no overlap on the device hides).
static bool slab_contiguous(int32_t count, gprb_dataset* const* ds, const double* const* X, int64_t ldx) {
  if (count < 1 || !ds[0]->slab || ds[0]->slab->count != count || ldx != ds[0]->d) return false;
  for (int i = 0; i < count; ++i) {
    if (ds[i]->slab != ds[0]->slab || ds[i]->slab_index != i) return false;
    if (X[i] != X[0] + (int64_t)i * ds[0]->n * ldx) return false;
  }
  return true;
}

static int slab_upload(gprb_ctx* ctx, int32_t count, gprb_dataset* const* ds, const double* X0) {
  gprb_slab* sl = ds[0]->slab;
  const int64_t n = ds[0]->n, npad = ds[0]->npad;
  const int d = ds[0]->d;
  for (int i = 0; i < count; ++i) ds[i]->version++;
  GPRB_CUDA(cudaMemcpyAsync(sl->X, X0, sizeof(double) * (size_t)count * n * d, cudaMemcpyHostToDevice, ctx->upload));
  int rc = launch_transpose_inputs_batched(sl->X, sl->Xt, (int)n, (int)npad, d, count, ctx->upload);
  if (rc) return rc;
  ctx->launches++;
  GPRB_CUDA(cudaStreamSynchronize(ctx->upload));
  return 0;
}

int gprb_datasets_create(gprb_ctx* ctx, int32_t count, int64_t n, int32_t d, const double* const* X, int64_t ldx,
                         gprb_dataset** out) {
  GPRB_REQUIRE(ctx && X && out && count >= 1, "gprb_datasets_create: bad argument");
  GPRB_REQUIRE(n >= 1 && n <= (1 << 20), "gprb_datasets_create: n out of range");
  GPRB_REQUIRE(d >= 1 && d <= MAX_D, "gprb_datasets_create: d must be in 1..62");
  GPRB_REQUIRE(ldx >= d, "gprb_datasets_create: ldx < d");
  for (int i = 0; i < count; ++i) { GPRB_REQUIRE(X[i] != nullptr, "gprb_datasets_create: NULL matrix"); out[i] = nullptr; }
  GPRB_CUDA(cudaSetDevice(ctx->device));
  const int64_t npad = (n + NB - 1) / NB * NB;
  gprb_slab* sl = new (std::nothrow) gprb_slab();
  GPRB_REQUIRE(sl != nullptr, "gprb_datasets_create: out of host memory");
  int rc;
  if ((rc = dev_alloc(&sl->X, (size_t)count * n * d)) || (rc = dev_alloc(&sl->Xt, (size_t)count * npad * d))) {
    cudaFree(sl->X); cudaFree(sl->Xt); delete sl;
    return rc;
  }
  sl->count = count; sl->refs = count;
  for (int i = 0; i < count; ++i) {
    gprb_dataset* ds = new (std::nothrow) gprb_dataset();
    if (!ds) {  // unwind
      for (int k = 0; k < i; ++k) { delete out[k]; out[k] = nullptr; }
      cudaFree(sl->X); cudaFree(sl->Xt); delete sl;
      set_error("gprb_datasets_create: out of host memory");
      return GPRB_ERR_ARG;
    }
    ds->ctx = ctx; ds->n = n; ds->d = d; ds->npad = npad;
    ds->X = sl->X + (size_t)i * n * d;
    ds->Xt = sl->Xt + (size_t)i * npad * d;
    ds->slab = sl; ds->slab_index = i;
    out[i] = ds;
  }
  rc = gprb_datasets_update(ctx, count, out, X, ldx);
  if (rc) {
    for (int i = 0; i < count; ++i) { delete out[i]; out[i] = nullptr; }
    cudaFree(sl->X); cudaFree(sl->Xt); delete sl;
  }
  return rc;
}

int gprb_datasets_update(gprb_ctx* ctx, int32_t count, gprb_dataset* const* ds, const double* const* X, int64_t ldx) {
  GPRB_REQUIRE(ctx && ds && X && count >= 0, "gprb_datasets_update: bad argument");
  for (int i = 0; i < count; ++i) {
    GPRB_REQUIRE(ds[i] && X[i], "gprb_datasets_update: NULL dataset or matrix");
    GPRB_REQUIRE(ds[i]->ctx == ctx, "gprb_datasets_update: dataset belongs to another context");
    GPRB_REQUIRE(ldx >= ds[i]->d, "gprb_datasets_update: ldx < d");
  }
  GPRB_CUDA(cudaSetDevice(ctx->device));
  if (slab_contiguous(count, ds, X, ldx)) return slab_upload(ctx, count, ds, X[0]);
  // The copies are queued back to back on the upload stream (truly asynchronous when the host matrices are page-locked);
  // the transposes run on a second stream, a quarter of the datasets at a time, behind an event - a transpose between
  // two copies on one stream would stall the copy engine for a kernel launch each time (2.6 -> 1.4 ms per 100 datasets).
  // One synchronisation at the end keeps the "synchronous on return" contract.
  const int chunk = std::max(1, (count + 3) / 4);
  for (int i0 = 0; i0 < count; i0 += chunk) {
    const int i1 = std::min(count, i0 + chunk);
    for (int i = i0; i < i1; ++i) {
      int rc = dataset_copy_async(ds[i], X[i], ldx, ctx->upload);
      if (rc) return rc;
    }
    GPRB_CUDA(cudaEventRecord(ctx->upload_ev, ctx->upload));
    GPRB_CUDA(cudaStreamWaitEvent(ctx->upload2, ctx->upload_ev, 0));
    for (int i = i0; i < i1; ++i) {
      int rc = launch_transpose_inputs(ds[i]->X, ds[i]->Xt, (int)ds[i]->n, (int)ds[i]->npad, ds[i]->d, ctx->upload2);
      if (rc) return rc;
      ctx->launches++;
    }
  }
  GPRB_CUDA(cudaStreamSynchronize(ctx->upload));
  GPRB_CUDA(cudaStreamSynchronize(ctx->upload2));
  return GPRB_OK;
}

int gprb_dataset_destroy(gprb_dataset* ds) {
  if (!ds) return GPRB_OK;
  if (ds->slab) {
    if (--ds->slab->refs == 0) {
      cudaSetDevice(ds->ctx->device);
      cudaFree(ds->slab->X); cudaFree(ds->slab->Xt);
      delete ds->slab;
    }
  } else {
    cudaFree(ds->X); cudaFree(ds->Xt);
  }
  delete ds;
  return GPRB_OK;
}

// ------------------------------------------------------------------------------------------------
static int upload_targets(gprb_batch* b, const double* ymm) {
  GPRB_CUDA(cudaMemcpy2DAsync(b->ymm, sizeof(double) * b->npad, ymm, sizeof(double) * b->n, sizeof(double) * b->n, b->B,
                              cudaMemcpyHostToDevice, b->stream[0]));
  GPRB_CUDA(cudaStreamSynchronize(b->stream[0]));
  b->state_ok.assign(b->B, 0);
  b->inv_ok.assign(b->B, 0);
  b->v_ok.assign(b->B, 0);
  return 0;
}

int gprb_batch_create(gprb_ctx* ctx, int32_t B, gprb_dataset* const* ds, const double* ymm, int32_t kernel_kind,
                      gprb_batch** out) {
  GPRB_REQUIRE(ctx && ds && ymm && out, "gprb_batch_create: NULL argument");
  *out = nullptr;
  GPRB_REQUIRE(B >= 1 && B <= 65535, "gprb_batch_create: B must be in 1..65535");
  GPRB_REQUIRE(kernel_kind >= 0 && kernel_kind <= 3, "gprb_batch_create: unknown kernel_kind");
  for (int i = 0; i < B; ++i) {
    GPRB_REQUIRE(ds[i] != nullptr, "gprb_batch_create: NULL dataset");
    GPRB_REQUIRE(ds[i]->ctx == ctx, "gprb_batch_create: dataset belongs to another context");
    GPRB_REQUIRE(ds[i]->n == ds[0]->n && ds[i]->d == ds[0]->d, "gprb_batch_create: datasets differ in n or d");
  }
  GPRB_CUDA(cudaSetDevice(ctx->device));
  gprb_batch* b = new (std::nothrow) gprb_batch();
  GPRB_REQUIRE(b != nullptr, "gprb_batch_create: out of host memory");
  b->ctx = ctx; b->B = B; b->n = ds[0]->n; b->d = ds[0]->d; b->P = b->d + 2; b->kind = kernel_kind;
  b->npad = ds[0]->npad; b->J = (int)(b->npad / NB);
  b->ds.assign(ds, ds + B);
  const size_t mat = (size_t)b->npad * b->npad, dinv = (size_t)b->J * NB * NB;
  const size_t ntiles = (size_t)b->J * (b->J + 1) / 2;
  int rc = 0;
  do {
    if ((rc = dev_alloc(&b->Xptr, B)) || (rc = dev_alloc(&b->Xtptr, B)) || (rc = dev_alloc(&b->ymm, (size_t)B * b->npad)) ||
        (rc = dev_alloc(&b->theta, (size_t)B * b->P)) || (rc = dev_alloc(&b->A, mat * B)) || (rc = dev_alloc(&b->Lm, mat * B)) ||
        (rc = dev_alloc(&b->Dinv, dinv * B)) || (rc = dev_alloc(&b->DinvT, dinv * B)) || (rc = dev_alloc(&b->KinvD, dinv * B)) ||
        (rc = dev_alloc(&b->alpha, (size_t)B * b->npad)) || (rc = dev_alloc(&b->zbuf, (size_t)B * b->npad)) ||
        (rc = dev_alloc(&b->jitter, B)) || (rc = dev_alloc(&b->diag_off, B)) || (rc = dev_alloc(&b->logdet_part, (size_t)B * b->J)) ||
        (rc = dev_alloc(&b->fail, B)) || (rc = dev_alloc(&b->mll, B)) || (rc = dev_alloc(&b->grad, (size_t)B * b->P)) ||
        (rc = dev_alloc(&b->grad_part, (size_t)B * ntiles * GRAD_PARTS_PER_TILE * b->P)) || (rc = dev_alloc(&b->list, 2 * (size_t)B)))
      break;
    b->stage_doubles = (int64_t)B * (b->P + 2);
    cudaError_t e;
    if ((e = cudaMallocHost((void**)&b->list_host, sizeof(int32_t) * 2 * B)) != cudaSuccess ||
        (e = cudaMallocHost((void**)&b->fail_host, sizeof(int32_t) * B)) != cudaSuccess ||
        (e = cudaMallocHost((void**)&b->stage_host, sizeof(double) * b->stage_doubles)) != cudaSuccess) {
      rc = cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__);
      break;
    }
    // operand tensor maps of the tile GEMM: padded boxes (132 / 68 rows) x KT columns x one GP
    if ((rc = make_tensor_map(&b->tm_L132, b->Lm, (uint64_t)b->npad, (uint64_t)b->npad, B, (uint64_t)b->npad, mat, NB + 4, KT)) ||
        (rc = make_tensor_map(&b->tm_L68, b->Lm, (uint64_t)b->npad, (uint64_t)b->npad, B, (uint64_t)b->npad, mat, NB / 2 + 4, KT)) ||
        (rc = make_tensor_map(&b->tm_A68, b->A, (uint64_t)b->npad, (uint64_t)b->npad, B, (uint64_t)b->npad, mat, NB / 2 + 4, KT)) ||
        (rc = make_tensor_map(&b->tm_DT132, b->DinvT, NB, (uint64_t)b->J * NB, B, NB, dinv, NB + 4, KT)) ||
        (rc = make_tensor_map(&b->tm_DT68, b->DinvT, NB, (uint64_t)b->J * NB, B, NB, dinv, NB / 2 + 4, KT)) ||
        (rc = make_tensor_map(&b->tm_D132, b->Dinv, NB, (uint64_t)b->J * NB, B, NB, dinv, NB + 4, KT)))
      break;
    b->nstreams = 4;
    b->rl_max = 24;
    if (const char* ev = getenv("GPRB200_RL_MAX")) b->rl_max = atoi(ev);
    // launches of the substitution with at most a quarter of the SM count in GPs (<= 37 on a B200: the reference's per-GP
    // call pattern, straggler rounds of the optimiser, the stream groups of the 8-GPU strong split) run one thread-block
    // cluster per GP; above that the one-CTA-per-GP kernel already streams at the HBM roof and clusters only add barriers
    // (measured: 100-GP groups, 122.0 -> 129.2 ms per 400-GP step with clusters)
    b->solve_cluster_below = ctx->sm_count / 4 + 1;
    b->group_skew = 0.0;
    if (const char* ev = getenv("GPRB200_GROUP_SKEW")) b->group_skew = std::max(0.0, std::min(0.9, atof(ev)));
    if (const char* ev = getenv("GPRB200_SOLVE_CLUSTER_BELOW")) b->solve_cluster_below = atoi(ev);
    if (const char* ev = getenv("GPRB200_STREAMS")) b->nstreams = std::max(1, std::min(MAX_STREAMS, atoi(ev)));
    for (int s = 0; s < MAX_STREAMS && !rc; ++s) {
      if ((e = cudaStreamCreateWithFlags(&b->stream[s], cudaStreamNonBlocking)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&b->join[s], cudaEventDisableTiming)) != cudaSuccess)
        rc = cuda_fail(e, "stream/event create", __FILE__, __LINE__);
    }
    for (int s = 0; s < 8 && !rc; ++s)
      if ((e = cudaEventCreate(&b->ev[s])) != cudaSuccess) rc = cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__);
    if (rc) break;
    std::vector<const double*> xp(B), xtp(B);
    for (int i = 0; i < B; ++i) { xp[i] = ds[i]->X; xtp[i] = ds[i]->Xt; }
    if ((e = cudaMemcpy(b->Xptr, xp.data(), sizeof(double*) * B, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(b->Xtptr, xtp.data(), sizeof(double*) * B, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemset(b->theta, 0, sizeof(double) * B * b->P)) != cudaSuccess ||
        (e = cudaMemset(b->jitter, 0, sizeof(double) * B)) != cudaSuccess ||
        (e = cudaMemset(b->diag_off, 0, sizeof(double) * B)) != cudaSuccess ||
        (e = cudaMemset(b->ymm, 0, sizeof(double) * b->npad * B)) != cudaSuccess ||  // zero padding, written once
        (e = cudaMemset(b->fail, 0, sizeof(int32_t) * B)) != cudaSuccess ||
        // k_diag_factor only writes the triangular halves of the inverted diagonal blocks: the zero halves are set here
        (e = cudaMemset(b->Dinv, 0, sizeof(double) * dinv * B)) != cudaSuccess ||
        (e = cudaMemset(b->DinvT, 0, sizeof(double) * dinv * B)) != cudaSuccess) {
      rc = cuda_fail(e, "batch init copies", __FILE__, __LINE__);
      break;
    }
    rc = upload_targets(b, ymm);
  } while (0);
  if (rc) { free_batch(b); return rc; }
  *out = b;
  return GPRB_OK;
}

int gprb_batch_set_targets(gprb_batch* b, const double* ymm) {
  GPRB_REQUIRE(b && ymm, "gprb_batch_set_targets: NULL argument");
  GPRB_CUDA(cudaSetDevice(b->ctx->device));
  return upload_targets(b, ymm);
}

int gprb_batch_set_diag_offset(gprb_batch* b, const double* offset) {
  GPRB_REQUIRE(b, "gprb_batch_set_diag_offset: NULL batch");
  GPRB_CUDA(cudaSetDevice(b->ctx->device));
  if (offset) {
    for (int i = 0; i < b->B; ++i) GPRB_REQUIRE(isfinite(offset[i]), "gprb_batch_set_diag_offset: non-finite offset");
    GPRB_CUDA(cudaMemcpy(b->diag_off, offset, sizeof(double) * b->B, cudaMemcpyHostToDevice));
  } else {
    GPRB_CUDA(cudaMemset(b->diag_off, 0, sizeof(double) * b->B));
  }
  // a different matrix is factorised from now on: nothing resident may be reused
  b->state_ok.assign(b->B, 0); b->inv_ok.assign(b->B, 0); b->v_ok.assign(b->B, 0);
  std::fill(b->theta_valid.begin(), b->theta_valid.end(), (uint8_t)0);
  return GPRB_OK;
}

int gprb_batch_destroy(gprb_batch* b) {
  if (b) { cudaSetDevice(b->ctx->device); cudaDeviceSynchronize(); }
  free_batch(b);
  return GPRB_OK;
}

int gprb_set_profiling(gprb_batch* b, int32_t on) {
  GPRB_REQUIRE(b, "gprb_set_profiling: NULL batch");
  b->profiling = on != 0;
  return GPRB_OK;
}

int gprb_last_gemm_launch_ms(gprb_batch* b, double* out, int32_t cap) {
  GPRB_REQUIRE(b && (out || cap == 0), "gprb_last_gemm_launch_ms: NULL argument");
  const int n = (int)b->gemm_ms.size();
  for (int i = 0; i < n && i < cap; ++i) out[i] = b->gemm_ms[i];
  return n;
}

int gprb_last_stage_ms(gprb_batch* b, double out[8]) {
  GPRB_REQUIRE(b && out, "gprb_last_stage_ms: NULL argument");
  for (int i = 0; i < 8; ++i) out[i] = b->stage_ms[i];
  return GPRB_OK;
}

// ------------------------------------------------------------------------------------------------
int gprb_eval_mixed(gprb_batch* b, const double* theta, const uint8_t* mode, double* mll, double* grad, int32_t* info) {
  GPRB_REQUIRE(b && theta && mode && mll && info, "gprb_eval_mixed: NULL argument");
  GPRB_CUDA(cudaSetDevice(b->ctx->device));
  bool any_grad = false;
  std::vector<int32_t> act;
  for (int i = 0; i < b->B; ++i) {
    GPRB_REQUIRE(mode[i] <= 2, "gprb_eval_mixed: mode must be 0 (skip), 1 (value) or 2 (value+gradient)");
    if (mode[i]) act.push_back(i);
    any_grad = any_grad || mode[i] == 2;
  }
  GPRB_REQUIRE(!any_grad || grad, "gprb_eval_mixed: a GP asks for its gradient but grad is NULL");
  int rc = upload_theta_rows(b, theta, act);
  if (rc) return rc;
  std::vector<int32_t> inf;
  if ((rc = evaluate_modes(b, theta, mode, inf))) return rc;
  return download_results(b, mode, inf, nullptr, mll, grad, info);
}

int gprb_eval(gprb_batch* b, const double* theta, const uint8_t* active, double* mll, double* grad, int32_t* info) {
  GPRB_REQUIRE(b && theta && mll && info, "gprb_eval: NULL argument");
  std::vector<uint8_t> mode(b->B);
  for (int i = 0; i < b->B; ++i) mode[i] = (!active || active[i]) ? (grad ? 2 : 1) : 0;
  return gprb_eval_mixed(b, theta, mode.data(), mll, grad, info);
}

int gprb_eval_device(gprb_batch* b, const double* theta_dev, double* mll_dev, double* grad_dev, int32_t* info_dev,
                     void* stream) {
  GPRB_REQUIRE(b && theta_dev && mll_dev, "gprb_eval_device: NULL argument");
  GPRB_CUDA(cudaSetDevice(b->ctx->device));
  cudaStream_t user = (cudaStream_t)stream;
  const int B = b->B, P = b->P;
  // order after the caller's stream, run on the batch streams, hand results back on the caller's stream
  GPRB_CUDA(cudaEventRecord(b->join[0], user));
  GPRB_CUDA(cudaStreamWaitEvent(b->stream[0], b->join[0], 0));
  GPRB_CUDA(cudaMemcpyAsync(b->theta, theta_dev, sizeof(double) * B * P, cudaMemcpyDeviceToDevice, b->stream[0]));
  std::vector<uint8_t> mode(B, grad_dev ? 2 : 1);
  std::vector<int32_t> inf;
  int rc = evaluate_modes(b, nullptr, mode.data(), inf);  // theta lives on the device: nothing to compare, nothing remembered
  if (rc) return rc;
  GPRB_CUDA(cudaMemcpyAsync(mll_dev, b->mll, sizeof(double) * B, cudaMemcpyDeviceToDevice, b->stream[0]));
  if (grad_dev) GPRB_CUDA(cudaMemcpyAsync(grad_dev, b->grad, sizeof(double) * B * P, cudaMemcpyDeviceToDevice, b->stream[0]));
  if (info_dev) {
    memcpy(b->fail_host, inf.data(), sizeof(int32_t) * B);
    GPRB_CUDA(cudaMemcpyAsync(info_dev, b->fail_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, b->stream[0]));
  }
  GPRB_CUDA(cudaEventRecord(b->join[0], b->stream[0]));
  GPRB_CUDA(cudaStreamWaitEvent(user, b->join[0], 0));
  GPRB_CUDA(cudaStreamSynchronize(b->stream[0]));
  return GPRB_OK;
}

// ------------------------------------------------------------------------------------------------
}  // extern "C"

template <typename T>
static int ensure_dev(T** p, size_t* cap, size_t need, bool zero = false) {
  if (*p && *cap >= need) return 0;
  cudaFree(*p);
  *p = nullptr; *cap = 0;
  int rc = dev_alloc(p, need);
  if (rc) return rc;
  if (zero) GPRB_CUDA(cudaMemset(*p, 0, sizeof(T) * need));
  *cap = need;
  return 0;
}

static int ensure_pinned(double** p, size_t* cap, size_t need) {
  if (*p && *cap >= need) return 0;
  if (*p) cudaFreeHost(*p);
  *p = nullptr; *cap = 0;
  cudaError_t e = cudaMallocHost((void**)p, sizeof(double) * std::max<size_t>(need, 1));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost(predict staging)", __FILE__, __LINE__);
  *cap = need;
  return 0;
}

// Enqueue the prediction of the GPs gp0 .. gp1-1 on pipeline `slot` (streams 4 slot .. 4 slot + 3): inputs are copied to
// the slot's pinned staging, everything else is asynchronous; predict_collect waits and hands the results out.
static int predict_enqueue(gprb_batch* b, int slot, int gp0, int gp1, int64_t m, const double* Xstar, int64_t xstar_stride,
                           const double* mstar, bool want_var, int gpb = 1) {
  gprb_predict_slot& sl = b->ps[slot];
  const int B = b->B, J = b->J, count = gp1 - gp0;
  cudaStream_t* str = b->stream + 4 * slot;
  cudaStream_t st = str[0];
  const int nblocks = (count + gpb - 1) / gpb;  // Xstar blocks: one per gpb consecutive GPs (the outputs of a trial)
  const size_t nx = (size_t)(xstar_stride ? xstar_stride * (nblocks - 1) + m * b->d : m * b->d);
  const size_t nout = (size_t)count * m;
  int rc;
  if (!sl.t0) {
    GPRB_CUDA(cudaEventCreate(&sl.t0));
    GPRB_CUDA(cudaEventCreate(&sl.t1));
    GPRB_CUDA(cudaMallocHost((void**)&sl.h_mask, sizeof(int32_t) * B));
    if ((rc = dev_alloc(&sl.mask, (size_t)B))) return rc;
  }
  if ((rc = ensure_dev(&sl.pX, &sl.pX_cap, nx)) || (rc = ensure_dev(&sl.pmu, &sl.pmu_cap, nout)) ||
      (rc = ensure_pinned(&sl.h_in, &sl.h_in_cap, nx + (mstar ? nout : 0))) ||
      (rc = ensure_pinned(&sl.h_out, &sl.h_out_cap, 2 * nout)))
    return rc;
  if (mstar && (rc = ensure_dev(&sl.pms, &sl.pms_cap, nout))) return rc;
  if (want_var && (rc = ensure_dev(&sl.pvar, &sl.pvar_cap, nout))) return rc;
  // per-GP mask: a GP without an evaluated state gets NaN rows and blocks nobody (examples/parallel/core.jl:41-46)
  bool all_v = true, any_ok = false;
  for (int i = 0; i < B; ++i) sl.h_mask[i] = b->state_ok[i] ? 0 : 1;
  for (int i = gp0; i < gp1; ++i)
    if (b->state_ok[i]) { any_ok = true; all_v = all_v && b->v_ok[i]; }
  memcpy(sl.h_in, Xstar, sizeof(double) * nx);
  if (mstar) memcpy(sl.h_in + nx, mstar, sizeof(double) * nout);
  GPRB_CUDA(cudaEventRecord(sl.t0, st));
  GPRB_CUDA(cudaMemcpyAsync(sl.mask, sl.h_mask, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));
  GPRB_CUDA(cudaMemcpyAsync(sl.pX, sl.h_in, sizeof(double) * nx, cudaMemcpyHostToDevice, st));
  if (mstar) GPRB_CUDA(cudaMemcpyAsync(sl.pms, sl.h_in + nx, sizeof(double) * nout, cudaMemcpyHostToDevice, st));

  // GEMV-like path: few test columns and the triangular inverse already resident (the last evaluation had a gradient)
  const int ps = m <= 1 ? 1 : m <= 2 ? 2 : m <= 4 ? 4 : 8;
  const size_t small_smem = ((size_t)ps * b->npad + ps * MAX_D + MAX_D + 8) * sizeof(double);
  if (want_var && m <= 8 && all_v && any_ok && small_smem <= 227 * 1024) {
    const int ncc = (int)((m + ps - 1) / ps);
    // CTAs sharing the row chunks of one (GP, column chunk): enough to put ~2 CTAs on every SM, a power of two <= 16
    int rs = 1;
    while (rs < PRED_CHUNKS && (int64_t)count * ncc * rs < 2 * (int64_t)b->ctx->sm_count) rs *= 2;
    if ((rc = ensure_dev(&sl.pq, &sl.pq_cap, (size_t)count * ncc * PRED_CHUNKS * 16)) ||
        (rc = ensure_dev(&sl.pcount, &sl.pcount_cap, (size_t)count * ncc, true)))
      return rc;
    PredictArgs pa{b->Xptr, b->theta, b->alpha, b->Lm, b->DinvT, sl.pX, mstar ? sl.pms : nullptr, sl.pmu, sl.pvar,
                   sl.pq, sl.pcount, sl.mask, xstar_stride, b->npad * b->npad, (int64_t)J * NB * NB,
                   (int)b->n, (int)b->npad, b->d, (int)m, b->kind, gp0, rs, gpb};
    if ((rc = launch_predict(pa, count, st))) return rc;
    b->ctx->launches++;
  } else {
    if ((rc = ensure_dev(&sl.pmupart, &sl.pmupart_cap, (size_t)count * J * PT))) return rc;
    if (want_var) {
      if ((rc = ensure_dev(&sl.pT, &sl.pT_cap, (size_t)count * b->npad * PT))) return rc;
      if (sl.tm_T_cap != sl.pT_cap) {  // (re)allocated: the right-hand-side block [GP][row r = k][PT test columns], box = 68 columns x KT rows
        if ((rc = make_tensor_map(&sl.tm_T68, sl.pT, PT, (uint64_t)b->npad, sl.pT_cap / ((size_t)b->npad * PT), PT, (uint64_t)b->npad * PT,
                                  NB / 2 + 4, KT)))
          return rc;
        sl.tm_T_cap = sl.pT_cap;
      }
    }
    const int nv = (int)((b->n + KT - 1) / KT * KT);
    for (int64_t s0 = 0; s0 < m; s0 += PT) {
      PredictTileArgs ta{b->Xtptr, b->theta, b->alpha, sl.pX, mstar ? sl.pms : nullptr, want_var ? sl.pT : nullptr, sl.pmupart,
                         sl.pmu, want_var ? sl.pvar : nullptr, xstar_stride, (int)b->n, (int)b->npad, b->d, J, (int)m, b->kind,
                         (int)s0, (int)std::min<int64_t>(PT, m - s0)};
      ta.gp_off = gp0;
      ta.mask = sl.mask;
      ta.gpb = gpb;
      { static const char* ev = getenv("GPRB200_PC_ZMAX"); if (ev) ta.zmax = atof(ev); }
      if ((rc = launch_predict_cross(ta, count, st))) return rc;
      b->ctx->launches++;
      if (want_var) {
        // L^-1 K*: one launch per block row, `count` tiles of 128 x 128 each (PDMats whiten! as a blocked substitution)
        GemmArgs ga{b->Lm, b->DinvT, b->Dinv, nullptr, b->Lm, b->KinvD, nullptr, b->npad * b->npad, (int64_t)J * NB * NB,
                    (int)b->npad, J, 0, GEMM_FWD_ROW, nv};
        ga.Tm = sl.pT; ga.t_stride = b->npad * PT; ga.ldt = PT; ga.t_gp_off = gp0;
        set_gemm_maps(ga, b);
        ga.tm_T68 = sl.tm_T68;
        ga.ncols = ta.mc;
        ga.fail = sl.mask;  // tiles of a GP without state exit at once
        // right-hand-side tiles are 64 test columns wide (two CTAs per SM); few GPs (large n, or the reference's single-GP
        // call pattern): 32-column tiles, so that a launch still puts a CTA on every SM
        ga.colw = (int64_t)count * ((ta.mc + 63) / 64) >= b->ctx->sm_count ? 64 : 32;
        const int ntl = (ta.mc + ga.colw - 1) / ga.colw;
        // the GPs are dealt over the stream groups: the dependent chain of J launches of one group fills the wave
        // tails of the others (`count` tiles per launch are not a multiple of the SM count)
        const int ns = std::min(4, b->nstreams);
        const int S = (count >= 8 * ns) ? ns : 1;
        if (S > 1) GPRB_CUDA(cudaEventRecord(b->join[4 * slot], st));
        for (int s = 0; s < S; ++s) {
          const int g0 = (int)((int64_t)count * s / S), g1 = (int)((int64_t)count * (s + 1) / S);
          cudaStream_t ss = str[s];
          if (s > 0) GPRB_CUDA(cudaStreamWaitEvent(ss, b->join[4 * slot], 0));
          ga.gp_off = gp0 + g0;
          for (int i = 0; i < J; ++i) {
            ga.step = i;
            if ((rc = launch_tile_gemm(ga, ntl, g1 - g0, ss))) return rc;
            b->ctx->launches++;
          }
          if (s > 0) {
            GPRB_CUDA(cudaEventRecord(b->join[4 * slot + s], ss));
            GPRB_CUDA(cudaStreamWaitEvent(st, b->join[4 * slot + s], 0));
          }
        }
      }
      if ((rc = launch_predict_finish(ta, count, st))) return rc;
      b->ctx->launches++;
    }
  }
  GPRB_CUDA(cudaMemcpyAsync(sl.h_out, sl.pmu, sizeof(double) * nout, cudaMemcpyDeviceToHost, st));
  if (want_var) GPRB_CUDA(cudaMemcpyAsync(sl.h_out + nout, sl.pvar, sizeof(double) * nout, cudaMemcpyDeviceToHost, st));
  GPRB_CUDA(cudaEventRecord(sl.t1, st));
  sl.pending = true; sl.gp0 = gp0; sl.gp1 = gp1; sl.m = m; sl.want_var = want_var;
  return 0;
}

static int predict_collect(gprb_batch* b, int slot, double* mu, double* var) {
  gprb_predict_slot& sl = b->ps[slot];
  GPRB_CUDA(cudaEventSynchronize(sl.t1));
  sl.pending = false;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, sl.t0, sl.t1) == cudaSuccess) sl.last_ms = ms;
  const size_t nout = (size_t)(sl.gp1 - sl.gp0) * sl.m;
  memcpy(mu, sl.h_out, sizeof(double) * nout);
  if (var && sl.want_var) memcpy(var, sl.h_out + nout, sizeof(double) * nout);
  return 0;
}

static int predict_check(gprb_batch* b, int slot, int gp0, int gp1, int64_t m, const double* Xstar, int64_t xstar_stride) {
  GPRB_REQUIRE(b && Xstar, "gprb_predict: NULL argument");
  GPRB_REQUIRE(slot == 0 || slot == 1, "gprb_predict: slot must be 0 or 1");
  GPRB_REQUIRE(gp0 >= 0 && gp0 < gp1 && gp1 <= b->B, "gprb_predict: bad GP range");
  GPRB_REQUIRE(m >= 1 && m <= (1 << 24), "gprb_predict: m out of range");
  GPRB_REQUIRE(xstar_stride == 0 || xstar_stride >= m * b->d, "gprb_predict: xstar_stride must be 0 or >= d*m");
  GPRB_REQUIRE(!b->ps[slot].pending, "gprb_predict: the slot has a prediction in flight - call gprb_predict_wait first");
  return 0;
}

extern "C" {

int gprb_predict(gprb_batch* b, int64_t m, const double* Xstar, int64_t xstar_stride, const double* mstar, double* mu,
                 double* var) {
  GPRB_REQUIRE(b && mu, "gprb_predict: NULL argument");
  int rc = predict_check(b, 0, 0, b->B, m, Xstar, xstar_stride);
  if (rc) return rc;
  GPRB_CUDA(cudaSetDevice(b->ctx->device));
  if ((rc = predict_enqueue(b, 0, 0, b->B, m, Xstar, xstar_stride, mstar, var != nullptr))) return rc;
  return predict_collect(b, 0, mu, var);
}

int gprb_predict_async(gprb_batch* b, int32_t slot, int32_t gp0, int32_t gp1, int64_t m, const double* Xstar,
                       int64_t xstar_stride, int32_t gps_per_block, const double* mstar, int32_t want_var) {
  int rc = predict_check(b, slot, gp0, gp1, m, Xstar, xstar_stride);
  if (rc) return rc;
  GPRB_REQUIRE(gps_per_block >= 1, "gprb_predict_async: gps_per_block must be >= 1");
  GPRB_CUDA(cudaSetDevice(b->ctx->device));
  return predict_enqueue(b, slot, gp0, gp1, m, Xstar, xstar_stride, mstar, want_var != 0, gps_per_block);
}

int gprb_predict_wait(gprb_batch* b, int32_t slot, double* mu, double* var) {
  GPRB_REQUIRE(b && mu && (slot == 0 || slot == 1), "gprb_predict_wait: bad argument");
  GPRB_REQUIRE(b->ps[slot].pending, "gprb_predict_wait: no prediction in flight on this slot");
  GPRB_REQUIRE(!var || b->ps[slot].want_var, "gprb_predict_wait: var requested but the prediction was enqueued with want_var = 0");
  GPRB_CUDA(cudaSetDevice(b->ctx->device));
  return predict_collect(b, slot, mu, var);
}

int gprb_last_predict_ms(gprb_batch* b, int32_t slot, double* ms) {
  GPRB_REQUIRE(b && ms && (slot == 0 || slot == 1), "gprb_last_predict_ms: bad argument");
  *ms = b->ps[slot].last_ms;
  return GPRB_OK;
}

// ------------------------------------------------------------------------------------------------
static int fetch_matrix(gprb_batch* b, const double* dev, std::vector<double>& host) {
  host.resize((size_t)b->npad * b->npad);
  GPRB_CUDA(cudaSetDevice(b->ctx->device));
  GPRB_CUDA(cudaStreamSynchronize(b->stream[0]));
  GPRB_CUDA(cudaMemcpy(host.data(), dev, sizeof(double) * host.size(), cudaMemcpyDeviceToHost));
  return 0;
}

int gprb_get_K(gprb_batch* b, int32_t gp, double* out) {
  GPRB_REQUIRE(b && out && gp >= 0 && gp < b->B, "gprb_get_K: bad argument");
  GPRB_REQUIRE(b->state_ok[gp], "gprb_get_K: no evaluated state");
  // K (noise and make_posdef! jitter included) stays resident in the lower tiles of A through every stage
  std::vector<double> h;
  int rc = fetch_matrix(b, b->A + (size_t)gp * b->npad * b->npad, h);
  if (rc) return rc;
  const int64_t n = b->n, np = b->npad;
  for (int64_t c = 0; c < n; ++c)
    for (int64_t r = c; r < n; ++r) out[r + c * n] = out[c + r * n] = h[r + c * np];
  return GPRB_OK;
}

int gprb_get_chol(gprb_batch* b, int32_t gp, double* out) {
  GPRB_REQUIRE(b && out && gp >= 0 && gp < b->B, "gprb_get_chol: bad argument");
  GPRB_REQUIRE(b->state_ok[gp], "gprb_get_chol: no evaluated state");
  std::vector<double> h;
  int rc = fetch_matrix(b, b->Lm + (size_t)gp * b->npad * b->npad, h);
  if (rc) return rc;
  const int64_t n = b->n, np = b->npad;
  for (int64_t c = 0; c < n; ++c)
    for (int64_t r = 0; r < n; ++r) out[r + c * n] = (r <= c) ? h[c + r * np] : 0.0;  // U(r,c) = L(c,r)
  return GPRB_OK;
}

int gprb_get_alpha(gprb_batch* b, int32_t gp, double* out) {
  GPRB_REQUIRE(b && out && gp >= 0 && gp < b->B, "gprb_get_alpha: bad argument");
  GPRB_REQUIRE(b->state_ok[gp], "gprb_get_alpha: no evaluated state");
  GPRB_CUDA(cudaSetDevice(b->ctx->device));
  GPRB_CUDA(cudaStreamSynchronize(b->stream[0]));
  GPRB_CUDA(cudaMemcpy(out, b->alpha + (size_t)gp * b->npad, sizeof(double) * b->n, cudaMemcpyDeviceToHost));
  return GPRB_OK;
}

int gprb_get_Kinv(gprb_batch* b, int32_t gp, double* out) {
  GPRB_REQUIRE(b && out && gp >= 0 && gp < b->B, "gprb_get_Kinv: bad argument");
  GPRB_REQUIRE(b->inv_ok[gp], "gprb_get_Kinv: the last evaluation was value-only (no inverse resident)");
  std::vector<double> h;
  int rc = fetch_matrix(b, b->A + (size_t)gp * b->npad * b->npad, h);
  if (rc) return rc;
  const size_t dn = (size_t)b->J * NB * NB;
  std::vector<double> dg(dn);
  GPRB_CUDA(cudaMemcpy(dg.data(), b->KinvD + (size_t)gp * dn, sizeof(double) * dn, cudaMemcpyDeviceToHost));
  const int64_t n = b->n, np = b->npad;
  for (int64_t c = 0; c < n; ++c)
    for (int64_t r = c; r < n; ++r) {
      const int64_t ti = r / NB, tj = c / NB, rl = r % NB, cl = c % NB;
      // tile (ti,tj), ti > tj, lives un-transposed at tile position (tj,ti); diagonal tiles in KinvD
      const double v = (ti == tj) ? dg[(size_t)ti * NB * NB + rl + cl * NB] : h[(tj * NB + rl) + (ti * NB + cl) * np];
      out[r + c * n] = out[c + r * n] = v;
    }
  return GPRB_OK;
}

}  // extern "C"
