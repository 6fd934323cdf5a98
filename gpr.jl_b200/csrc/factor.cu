// Diagonal-block factorisation (potf2 + trtri of one 128x128 block, in shared memory) and the
// triangular solves for alpha = K^-1 (y - m) with the log-marginal-likelihood value
// (SURVEY.md section 8a rows a6, a7).  The O(n^3) work lives in tilegemm.cu; these are the serial
// O(n * 128^2) / O(n^2) pieces between the GEMM launches.
#include "common.cuh"
#include "kernels.h"

namespace gprb {

constexpr int DIAG_THREADS = 256;
constexpr int LDD = NB + 1;  // padded column stride of the smem block

// One CTA per GP: Lm(j,j) holds S = K(j,j) - sum_k L(j,k) L(j,k)^T on entry (lower part valid).
// On exit: Lm(j,j) = L_jj (zeros above the diagonal), Dinv[j] = inv(L_jj), DinvT[j] = its transpose,
// logdet_part[j] = sum log diag(L_jj), fail = LAPACK-style info (first non-positive / NaN pivot, 1-based).
__global__ void __launch_bounds__(DIAG_THREADS, 1) k_diag_factor(DiagArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* S = reinterpret_cast<double*>(smem_raw);  // S(r,c) = S[r + c*LDD]
  double* diagW = S + NB * LDD;                      // 1 / L(c,c)
  __shared__ int bad_col;
  const int gp = g.list ? g.list[blockIdx.x] : blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = g.step;
  const int64_t npad = g.npad;
  if (g.fail[gp] != 0) return;  // already failed (or non-finite theta): results are discarded by the host
  double* T = g.Lm + (int64_t)gp * g.mat_stride + (int64_t)j * NB + (int64_t)j * NB * npad;
  for (int idx = tid; idx < NB * NB; idx += DIAG_THREADS) {
    const int r = idx & (NB - 1), c = idx >> 7;
    S[r + c * LDD] = T[r + c * npad];
  }
  if (tid == 0) bad_col = 0;

  // ---- unblocked right-looking Cholesky (lower), LAPACK dpotf2 failure rule: pivot <= 0 or NaN
  for (int c = 0; c < NB; ++c) {
    __syncthreads();
    double piv = S[c + c * LDD];
    if (!(piv > 0.0)) {
      if (tid == 0 && bad_col == 0) bad_col = c + 1;
      piv = 1.0;
    }
    const double l = sqrt(piv);
    const double inv = 1.0 / l;
    __syncthreads();
    for (int r = c + 1 + tid; r < NB; r += DIAG_THREADS) S[r + c * LDD] *= inv;
    if (tid == 0) S[c + c * LDD] = l;
    __syncthreads();
    for (int k = c + 1 + warp; k < NB; k += DIAG_THREADS / 32) {
      const double lk = S[k + c * LDD];
      for (int r = k + lane; r < NB; r += 32) S[r + k * LDD] = fma(-S[r + c * LDD], lk, S[r + k * LDD]);
    }
  }
  __syncthreads();
  if (bad_col != 0) {
    if (tid == 0) g.fail[gp] = j * NB + bad_col;
    return;
  }

  // ---- W = inv(L): two threads per column c (k-parity split), W(r,c) r>c parked at S(c,r) (strict upper)
  {
    const int c = tid >> 1, half = tid & 1;
    const double wcc = 1.0 / S[c + c * LDD];
    if (half == 0) diagW[c] = wcc;
    // warp-uniform row loop (lanes of a warp own columns 16*warp .. 16*warp+15), predicated on r > c
    for (int r = 16 * warp + 1; r < NB; ++r) {
      // s = sum_{k=c}^{r-1} L(r,k) W(k,c)
      double s0 = 0.0, s1 = 0.0;
      if (r > c) {
        int k = c + half;
        if (half == 0) { s0 = S[r + c * LDD] * wcc; k += 2; }
        for (; k + 2 < r; k += 4) {
          s0 = fma(S[r + k * LDD], S[c + k * LDD], s0);
          s1 = fma(S[r + (k + 2) * LDD], S[c + (k + 2) * LDD], s1);
        }
        for (; k < r; k += 2) s0 = fma(S[r + k * LDD], S[c + k * LDD], s0);
      }
      double s = s0 + s1;
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (r > c && half == 0) S[c + r * LDD] = -s / S[r + r * LDD];
      __syncwarp();
    }
  }
  __syncthreads();

  double* Dinv = g.Dinv + (int64_t)gp * g.dinv_stride + (int64_t)j * NB * NB;
  double* DinvT = g.DinvT + (int64_t)gp * g.dinv_stride + (int64_t)j * NB * NB;
  for (int idx = tid; idx < NB * NB; idx += DIAG_THREADS) {
    const int r = idx & (NB - 1), c = idx >> 7;
    const double lrc = S[r + c * LDD];           // L(r,c) if r >= c, W(c,r)... careful below
    // element (r,c): lower part of S holds L, strict upper part holds W transposed (S(c',r') = W(r',c'))
    double Lval, Wval, WTval;
    if (r > c) {
      Lval = lrc;                 // L(r,c)
      Wval = S[c + r * LDD];      // W(r,c) parked at S(c,r)
      WTval = 0.0;                // W^T(r,c) = W(c,r) = 0 (c < r)
    } else if (r == c) {
      Lval = lrc;
      Wval = diagW[c];
      WTval = diagW[c];
    } else {
      Lval = 0.0;
      Wval = 0.0;
      WTval = lrc;                // W^T(r,c) = W(c,r), c > r, parked at S(r,c)
    }
    T[r + c * npad] = Lval;
    Dinv[idx] = Wval;
    DinvT[idx] = WTval;
  }
  if (warp == 0) {
    double s = 0.0;
    for (int r = lane; r < NB; r += 32) s += log(S[r + r * LDD]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) g.logdet_part[(int64_t)gp * g.J + j] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// alpha = L^-T L^-1 ymm by blocked substitution with the inverted diagonal blocks; one CTA per GP.
// ---------------------------------------------------------------------------------------------
constexpr int SOLVE_THREADS = 512;

__global__ void __launch_bounds__(SOLVE_THREADS) k_solve(SolveArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* z = reinterpret_cast<double*>(smem_raw);  // [npad]
  double* al = z + g.npad;                            // [npad]
  double* red = al + g.npad;                          // [4][NB]
  double* rv = red + 4 * NB;                          // [NB]
  const int gp = g.list ? g.list[blockIdx.x] : blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t npad = g.npad;
  const int J = g.J;
  if (g.fail[gp] != 0) {
    if (tid == 0) g.mll[gp] = -__longlong_as_double(0x7ff0000000000000LL);
    return;
  }
  const double* L = g.Lm + (int64_t)gp * g.mat_stride;
  const double* Dinv = g.Dinv + (int64_t)gp * g.dinv_stride;
  const double* y = g.ymm + (int64_t)gp * npad;
  const int row = tid & (NB - 1), part = tid >> 7;  // 4 k-partitions

  // ---- forward: z_j = Dinv_j (y_j - sum_{k<j} L(j,k) z_k)
  for (int jb = 0; jb < J; ++jb) {
    const double* Lrow = L + (int64_t)jb * NB + row;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const int kend = jb * NB;
    int k = part;
    for (; k + 12 < kend; k += 16) {
      a0 = fma(Lrow[(int64_t)k * npad], z[k], a0);
      a1 = fma(Lrow[(int64_t)(k + 4) * npad], z[k + 4], a1);
      a2 = fma(Lrow[(int64_t)(k + 8) * npad], z[k + 8], a2);
      a3 = fma(Lrow[(int64_t)(k + 12) * npad], z[k + 12], a3);
    }
    for (; k < kend; k += 4) a0 = fma(Lrow[(int64_t)k * npad], z[k], a0);
    red[part * NB + row] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (tid < NB) rv[tid] = y[jb * NB + tid] - ((red[tid] + red[NB + tid]) + (red[2 * NB + tid] + red[3 * NB + tid]));
    __syncthreads();
    const double* Dj = Dinv + (int64_t)jb * NB * NB;
    double s = 0.0;
    for (int kk = part; kk <= row; kk += 4) s = fma(Dj[row + kk * NB], rv[kk], s);
    red[part * NB + row] = s;
    __syncthreads();
    if (tid < NB) z[jb * NB + tid] = (red[tid] + red[NB + tid]) + (red[2 * NB + tid] + red[3 * NB + tid]);
    __syncthreads();
  }

  // ---- backward: alpha_j = Dinv_j^T (z_j - sum_{i>j} L(i,j)^T alpha_i); warp per column, lanes over rows
  for (int jb = J - 1; jb >= 0; --jb) {
    const int rbeg = (jb + 1) * NB;
    for (int c = warp; c < NB; c += SOLVE_THREADS / 32) {
      const double* Lcol = L + (int64_t)(jb * NB + c) * npad;
      double a0 = 0.0, a1 = 0.0;
      int r = rbeg + lane;
      for (; r + 32 < (int)npad; r += 64) {
        a0 = fma(Lcol[r], al[r], a0);
        a1 = fma(Lcol[r + 32], al[r + 32], a1);
      }
      for (; r < (int)npad; r += 32) a0 = fma(Lcol[r], al[r], a0);
      double s = a0 + a1;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (lane == 0) rv[c] = z[jb * NB + c] - s;
    }
    __syncthreads();
    const double* Dj = Dinv + (int64_t)jb * NB * NB;
    for (int c = warp; c < NB; c += SOLVE_THREADS / 32) {
      double s = 0.0;
      for (int kk = c + lane; kk < NB; kk += 32) s = fma(Dj[kk + c * NB], rv[kk], s);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (lane == 0) al[jb * NB + c] = s;
    }
    __syncthreads();
  }

  double* zout = g.zbuf + (int64_t)gp * npad;
  double* aout = g.alpha + (int64_t)gp * npad;
  for (int r = tid; r < (int)npad; r += SOLVE_THREADS) { zout[r] = z[r]; aout[r] = al[r]; }

  // ---- mll = -1/2 (z'z + 2 sum log L_ii + n log 2pi)      (ymm' alpha == z'z)
  if (warp == 0) {
    double s = 0.0;
    for (int r = lane; r < g.n; r += 32) s = fma(z[r], z[r], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) {
      double ld = 0.0;
      for (int jb = 0; jb < J; ++jb) ld += g.logdet_part[(int64_t)gp * J + jb];
      g.mll[gp] = -0.5 * (s + 2.0 * ld + g.n * 1.8378770664093453);
    }
  }
}

int launch_diag_factor(const DiagArgs& a, int count, cudaStream_t stream) {
  if (count <= 0) return 0;
  const size_t smem = (size_t)(NB * LDD + NB) * sizeof(double);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_diag_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_diag_factor)", __FILE__, __LINE__);
    configured = true;
  }
  k_diag_factor<<<count, DIAG_THREADS, smem, stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_diag_factor launch", __FILE__, __LINE__);
  return 0;
}

int launch_solve(const SolveArgs& a, int count, cudaStream_t stream) {
  if (count <= 0) return 0;
  const size_t smem = (size_t)(2 * a.npad + 5 * NB) * sizeof(double);
  if (smem > 227 * 1024) {
    set_error("k_solve: n too large for the shared-memory resident substitution (npad <= 14000)");
    return GPRB_ERR_ARG;
  }
  cudaError_t e = cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_solve)", __FILE__, __LINE__);
  k_solve<<<count, SOLVE_THREADS, smem, stream>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_solve launch", __FILE__, __LINE__);
  return 0;
}

}  // namespace gprb
