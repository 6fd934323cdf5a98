// Diagonal-block factorisation (potf2 + trtri of one 128x128 block, in shared memory) and the
// triangular solves for alpha = K^-1 (y - m) with the log-marginal-likelihood value
// (SURVEY.md section 8a rows a6, a7).  The O(n^3) work lives in tilegemm.cu; these are the serial
// O(n * 128^2) / O(n^2) pieces between the GEMM launches.
#include <cooperative_groups.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace gprb {

constexpr int DIAG_THREADS = 256;
constexpr int SB = 32;            // sub-block edge inside the 128x128 diagonal block
constexpr int NSB = NB / SB;      // 4
constexpr int LDW = SB + 4;       // padded column stride of the 32x32 side buffers (conflict-free DMMA fragments)
constexpr unsigned FULL = 0xffffffffu;

// acc(8 x 32 slab) += A(8 x 4*ksteps) * B(32 x 4*ksteps)^T on the DMMA pipe, operands "k-major" in smem:
// A(m,k) = Abase[k*lda + m], B(n,k) = Bbase[k*ldb + n];  lane = 4g+t owns acc[ni] = C(g, ni*8 + 2t + {0,1}).
__device__ __forceinline__ void slab_mma(double (&acc)[4][2], const double* Abase, int lda, const double* Bbase, int ldb,
                                         int ksteps, int g, int t) {
  for (int k4 = 0; k4 < ksteps; ++k4) {
    const double a = Abase[(k4 * 4 + t) * lda + g];
    const double* bp = Bbase + (k4 * 4 + t) * ldb + g;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) dmma884(acc[ni][0], acc[ni][1], a, bp[ni * 8]);
  }
}

// Same product with the B operand stored the other way round: B(n,k) = Bbase[n*ldb + k] (a column-major sub-block read as
// its own transpose).  Bank pattern of a half-warp: (4g + t) mod 16 with ldb = 36 - conflict-free like the k-major form.
__device__ __forceinline__ void slab_mma_bt(double (&acc)[4][2], const double* Abase, int lda, const double* Bbase, int ldb,
                                            int ksteps, int g, int t) {
  for (int k4 = 0; k4 < ksteps; ++k4) {
    const double a = Abase[(k4 * 4 + t) * lda + g];
    const double* bp = Bbase + g * ldb + k4 * 4 + t;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) dmma884(acc[ni][0], acc[ni][1], a, bp[ni * 8 * ldb]);
  }
}

// Packed lower-triangular storage of the 128x128 block in shared memory: the ten 32x32 sub-blocks (bi >= bj), each
// column-major with the padded stride LDW = 36.  90 KB per CTA instead of the 195 KB of a padded square plus side
// buffers: TWO blocks are factorised per SM at a time (the serial 32x32 potf2 + trtri of one overlaps the DMMA phases of
// the other), and a factorisation can share an SM with a k_tile_gemm CTA (110.7 KB) of another stream group instead of
// waiting for both of its CTAs to drain.
constexpr int SBLK = SB * LDW;                     // doubles of one sub-block slot
constexpr int NSLOT = NSB * (NSB + 1) / 2;         // 10
__device__ __forceinline__ int slot_of(int bi, int bj) { return (bi * (bi + 1) / 2 + bj) * SBLK; }

// One CTA per GP.  On entry Lm(j,j) holds S = K(j,j) - sum_k L(j,k) L(j,k)^T (lower part valid).
// On exit: Lm(j,j) = L_jj, Dinv[j] = inv(L_jj), DinvT[j] = inv(L_jj)^T, logdet_part[j] = sum log diag(L_jj),
// fail = LAPACK-style info (first non-positive / NaN pivot, 1-based) when the block is not positive definite.
//
// Blocked inside shared memory with 32x32 sub-blocks: the sub-block on the diagonal is factorised and inverted by
// ONE warp entirely in registers (row per lane, pivots broadcast by shuffles); panel solves, trailing updates and
// the blocked triangular inverse are 8x32 DMMA slabs spread over the 8 warps.  Everything is in place: slot (i,j), i > j,
// holds S(i,j), then L(i,j) (written to HBM at the end of phase A), then W(i,j) = inv(L)(i,j) stored ROW-major (the B operand
// of the later block rows); slot (j,j) holds S(j,j), then W(j,j) column-major (L(j,j) goes to HBM from registers).
__global__ void __launch_bounds__(DIAG_THREADS, 2) k_diag_factor(DiagArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* S = reinterpret_cast<double*>(smem_raw);   // [NSLOT][SBLK]
  __shared__ int bad_col;
  __shared__ double dinv32[SB];
  const int gp = g.list ? g.list[blockIdx.x] : blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const int warp_u = __shfl_sync(FULL, warp, 0);  // warp index through the uniform datapath
  const int j = g.step;
  const int64_t npad = g.npad;
  if (g.fail[gp] != 0) return;  // already failed (or non-finite theta): results are discarded by the host
  double* T = g.Lm + (int64_t)gp * g.mat_stride + (int64_t)j * NB + (int64_t)j * NB * npad;
  const double* Tsrc = g.Src ? g.Src + (int64_t)gp * g.mat_stride + (int64_t)j * NB + (int64_t)j * NB * npad : T;
  // rows / cols beyond nv are padding the GEMM stages neither compute nor read: treat them as the identity here
  // (nv is a multiple of 16: a pair of rows (r, r + 1), r even, is valid or padding as a whole)
  const int nvl = (j == g.J - 1) ? g.nv - j * NB : NB;
  {
    // lower sub-blocks only, 16-byte loads, all of a thread's loads in flight at once (the read is pure latency otherwise)
    constexpr int PAIRS = NSLOT * SB * SB / 2, PER_THREAD = PAIRS / DIAG_THREADS;  // 5120 pairs, 20 per thread
    constexpr int BATCH = PER_THREAD / 2;  // two batches of ten loads: 40 data registers in flight
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double2 v[BATCH];
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const int idx = tid + (h * BATCH + u) * DIAG_THREADS, p = idx >> 9, w = idx & 511;  // 512 row pairs per sub-block
        const int bi = p < 1 ? 0 : p < 3 ? 1 : p < 6 ? 2 : 3, bj = p - bi * (bi + 1) / 2;
        const int r = bi * SB + 2 * (w & 15), c = bj * SB + (w >> 4);
        v[u] = (r < nvl && c < nvl) ? *reinterpret_cast<const double2*>(Tsrc + r + c * npad)
                                    : make_double2(r == c ? 1.0 : 0.0, r + 1 == c ? 1.0 : 0.0);
      }
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const int idx = tid + (h * BATCH + u) * DIAG_THREADS, p = idx >> 9, w = idx & 511;
        *reinterpret_cast<double2*>(&S[p * SBLK + (w >> 4) * LDW + 2 * (w & 15)]) = v[u];
      }
    }
  }
  if (tid == 0) bad_col = 0;
  __syncthreads();
  double logdet = 0.0;

  // ================= phase A: right-looking Cholesky over 32-wide sub-block columns =================
#pragma unroll 1
  for (int jj = 0; jj < NSB; ++jj) {  // not unrolled: the straight-line sub-block code below is reused by all four iterations
    const int j0 = jj * SB;
    double* D = S + slot_of(jj, jj);  // D(r,c) = D[c*LDW + r]
    if (warp_u == 0) {  // provably warp-uniform branch: the shuffles below compile to plain SHFL (no collective wrappers)
      // ---- A1: potf2 + trtri of the 32x32 diagonal sub-block by one warp, entirely in registers.
      // This is the serial critical path of the whole block (7 warps wait), so it is written for latency:
      //   potf2  right-looking, lane r owns row r (32 doubles); pivots / column entries broadcast by shuffles;
      //          one rsqrt per pivot (no sqrt + divide), logs taken once per lane afterwards
      //   trtri  column-oriented forward substitution, lane c owns column c of W = inv(L_dd); L(r,k) is a
      //          broadcast shared-memory read, all updates of one step are independent FMAs
      // Both are straight-line (fully unrolled, ~3.5k instructions, reused by the four sub-block iterations).
      double a[SB];
#pragma unroll
      for (int c = 0; c < SB; ++c) a[c] = D[c * LDW + lane];
      int bad = 0;
      double mydinv = 1.0;  // lane c ends up with 1 / L_cc
      // Software-pipelined over the columns: as soon as column c + 1 has its update from column c, its pivot is broadcast
      // and the reciprocal square root started, BEFORE the remaining trailing updates of column c are issued.  The warp
      // issues in order: with the pivot taken at the top of the next iteration, the ~120-cycle rsqrt could only start
      // after all 3 (31 - c) update instructions had been issued.  Same operations on the same operands: bit-identical.
      double piv = __shfl_sync(FULL, a[0], 0);
      if (!(piv > 0.0)) {  // LAPACK dpotf2 rule: pivot <= 0 or NaN
        bad = j0 + 1;
        piv = 1.0;
      }
      double inv = rsqrt(piv);
#pragma unroll
      for (int c = 0; c < SB; ++c) {
        const double lrc = (lane == c) ? piv * inv : a[c] * inv;  // L(lane, c), valid for lane >= c
        a[c] = lrc;
        if (lane == c) mydinv = inv;
        if (c + 1 < SB) {
          a[c + 1] = fma(-lrc, __shfl_sync(FULL, lrc, c + 1), a[c + 1]);
          piv = __shfl_sync(FULL, a[c + 1], c + 1);
          if (!(piv > 0.0)) {
            if (bad == 0) bad = j0 + c + 2;
            piv = 1.0;
          }
          inv = rsqrt(piv);
#pragma unroll
          for (int k = c + 2; k < SB; ++k) a[k] = fma(-lrc, __shfl_sync(FULL, lrc, k), a[k]);  // valid for lane >= k
        }
      }
      if (bad != 0 && lane == 0 && bad_col == 0) bad_col = bad;
      logdet -= log(mydinv);  // log L_cc = -log(1 / L_cc) of this lane's pivot; warp-reduced once at the end of the kernel
      // L_dd: lower part back into D (read by the inverse below) and straight to HBM, zeros above the diagonal
#pragma unroll
      for (int c = 0; c < SB; ++c) {
        const double v = (lane >= c) ? a[c] : 0.0;
        D[c * LDW + lane] = v;
        T[(j0 + lane) + (int64_t)(j0 + c) * npad] = v;
      }
      dinv32[lane] = mydinv;
      __syncwarp();
      // W = inv(L_dd), column `lane`: solve L w = e_lane
      double w[SB];
#pragma unroll
      for (int r = 0; r < SB; ++r) w[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < SB; ++k) {
        w[k] *= dinv32[k];
#pragma unroll
        for (int r = k + 1; r < SB; ++r) w[r] = fma(-D[k * LDW + r], w[k], w[r]);  // L(r,k): broadcast read
      }
      __syncwarp();
      // the slot becomes W_dd, column-major (exact zeros above the diagonal)
#pragma unroll
      for (int r = 0; r < SB; ++r) D[r + lane * LDW] = w[r];
    }
    __syncthreads();
    // ---- A2: panel  L(i, jj) = S(i, jj) * W_dd^T  for the sub-blocks below; 8-row slabs over the warps
    // (slab `sl` of sub-block `bi` is read and written by the same warp only: in place)
    const int nslab_panel = (NSB - 1 - jj) * (SB / 8);
    for (int it = warp; it < nslab_panel; it += DIAG_THREADS / 32) {
      const int sl = it & 3, bi = jj + 1 + (it >> 2);
      double* P = S + slot_of(bi, jj) + sl * 8;
      double acc[4][2] = {};
      slab_mma(acc, P, LDW, D, LDW, SB / 4, gq, t);
      __syncwarp();
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        P[(ni * 8 + 2 * t) * LDW + gq] = acc[ni][0];
        P[(ni * 8 + 2 * t + 1) * LDW + gq] = acc[ni][1];
      }
    }
    __syncthreads();
    // ---- A3: trailing update  S(bi, bk) -= L(bi, jj) L(bk, jj)^T  for jj < bk <= bi
    const int nrem = NSB - 1 - jj;
    const int nitems = nrem * (nrem + 1) / 2 * (SB / 8);
    for (int it = warp; it < nitems; it += DIAG_THREADS / 32) {
      const int sl = it & 3, blk = it >> 2;
      int bi = 0;
      while ((bi + 1) * (bi + 2) / 2 <= blk) ++bi;
      const int bk = blk - bi * (bi + 1) / 2;
      double acc[4][2] = {};
      slab_mma(acc, S + slot_of(jj + 1 + bi, jj) + sl * 8, LDW, S + slot_of(jj + 1 + bk, jj), LDW, SB / 4, gq, t);
      double* C = S + slot_of(jj + 1 + bi, jj + 1 + bk) + sl * 8;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        C[(ni * 8 + 2 * t) * LDW + gq] -= acc[ni][0];
        C[(ni * 8 + 2 * t + 1) * LDW + gq] -= acc[ni][1];
      }
    }
    __syncthreads();
  }
  if (bad_col != 0) {
    if (tid == 0) g.fail[gp] = j * NB + bad_col;
    return;
  }
  // L's off-diagonal sub-blocks go to HBM now (phase B overwrites them), zeros into the upper sub-blocks of the tile
  {
    constexpr int NOFF = NSLOT - NSB;  // 6
#pragma unroll
    for (int u = 0; u < NOFF * SB * SB / 2 / DIAG_THREADS; ++u) {  // 12 row pairs per thread
      const int idx = tid + u * DIAG_THREADS, q = idx >> 9, w = idx & 511;
      const int bi = q < 1 ? 1 : q < 3 ? 2 : 3, bj = q - (bi * (bi - 1)) / 2;
      const int rr = 2 * (w & 15), cc = w >> 4;
      const double2 v = *reinterpret_cast<const double2*>(&S[slot_of(bi, bj) + cc * LDW + rr]);
      *reinterpret_cast<double2*>(T + (bi * SB + rr) + (int64_t)(bj * SB + cc) * npad) = v;
      *reinterpret_cast<double2*>(T + (bj * SB + rr) + (int64_t)(bi * SB + cc) * npad) = make_double2(0.0, 0.0);
    }
  }

  // ================= phase B: blocked inverse, W(i,j) = -W_ii * sum_{k=j}^{i-1} L(i,k) W(k,j) =========================
  // At most 12 (slab, j) items per block row over the 8 warps: every warp keeps its (up to two) 8x32 results in registers
  // across the barrier that separates the last read of a slot from its overwrite.
  for (int i = 1; i < NSB; ++i) {
    const int nit = i * (SB / 8);
    double acc[2][4][2];
    // B1: T_j (32x32) = L(i, j..i-1) * W(j..i-1, j); the k = j term reads the column-major diagonal slot transposed
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int it = warp + u * (DIAG_THREADS / 32);
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) acc[u][ni][0] = acc[u][ni][1] = 0.0;
      if (it < nit) {
        const int sl = it & 3, jb = it >> 2;
        slab_mma_bt(acc[u], S + slot_of(i, jb) + sl * 8, LDW, S + slot_of(jb, jb), LDW, SB / 4, gq, t);
        for (int k = jb + 1; k < i; ++k) slab_mma(acc[u], S + slot_of(i, k) + sl * 8, LDW, S + slot_of(k, jb), LDW, SB / 4, gq, t);
      }
    }
    __syncthreads();  // every L(i, .) has been read: the slots of block row i now take T, stored T(k,n) at [k*LDW + n]
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int it = warp + u * (DIAG_THREADS / 32);
      if (it < nit) {
        const int sl = it & 3, jb = it >> 2;
        double* Tj = S + slot_of(i, jb);
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          *reinterpret_cast<double2*>(&Tj[(sl * 8 + gq) * LDW + ni * 8 + 2 * t]) = make_double2(acc[u][ni][0], acc[u][ni][1]);
      }
    }
    __syncthreads();
    // B2: W(i,j) = -W_ii * T_j (every slab needs the whole T_j: results wait in registers for the barrier again)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int it = warp + u * (DIAG_THREADS / 32);
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) acc[u][ni][0] = acc[u][ni][1] = 0.0;
      if (it < nit) {
        const int sl = it & 3, jb = it >> 2;
        slab_mma(acc[u], S + slot_of(i, i) + sl * 8, LDW, S + slot_of(i, jb), LDW, SB / 4, gq, t);
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int it = warp + u * (DIAG_THREADS / 32);
      if (it < nit) {
        const int sl = it & 3, jb = it >> 2;
        double* Wj = S + slot_of(i, jb);  // W(i,j)(r,c) at [r*LDW + c]
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          *reinterpret_cast<double2*>(&Wj[(sl * 8 + gq) * LDW + ni * 8 + 2 * t]) = make_double2(-acc[u][ni][0], -acc[u][ni][1]);
      }
    }
    __syncthreads();
  }

  // ================= write-out: W = inv(L_jj) column-major into Dinv, its transpose into DinvT =================
  double* Dinv = g.Dinv + (int64_t)gp * g.dinv_stride + (int64_t)j * NB * NB;
  double* DinvT = g.DinvT + (int64_t)gp * g.dinv_stride + (int64_t)j * NB * NB;
  // Only the ten lower sub-blocks of W (and the ten upper ones of W^T) are written: the other halves of both buffers are
  // structural zeros, set once when the batch is created and never touched again.
  // A thread transposes 2x2 entries: two 16-byte shared-memory loads along the slot's contiguous direction give it
  // W(r..r+1, c..c+1), which is one 16-byte store per output column of W and one per output row of W^T.  The lane layout
  // (4 pairs along the contiguous direction x 8 pairs across it) keeps the loads conflict-free with the 36-double pitch;
  // reading row pairs of a row-major slot with scalar loads, as a column-major output order suggests, was an 8-way conflict.
  for (int pc = warp; pc < NSLOT * 8; pc += DIAG_THREADS / 32) {  // 8 pieces of 16 x 8 (or 8 x 16) entries per slot
    const int q = pc >> 3, sub = pc & 7;
    const int bi = q < 1 ? 0 : q < 3 ? 1 : q < 6 ? 2 : 3, bj = q - bi * (bi + 1) / 2;
    const double* P = S + q * SBLK;
    double2 w0, w1, t0, t1;  // w0 / w1 = W(r..r+1, c) / W(r..r+1, c+1);  t0 / t1 = W(r, c..c+1) / W(r+1, c..c+1)
    int rr, cc;
    if (bi == bj) {  // column-major slot, W(r,c) at [c*LDW + r]: lanes 4 row pairs x 8 column pairs, piece = 8 rows x 16 columns
      rr = (sub & 3) * 8 + 2 * (lane & 3);
      cc = (sub >> 2) * 16 + 2 * (lane >> 2);
      w0 = *reinterpret_cast<const double2*>(&P[cc * LDW + rr]);
      w1 = *reinterpret_cast<const double2*>(&P[(cc + 1) * LDW + rr]);
      t0 = make_double2(w0.x, w1.x);
      t1 = make_double2(w0.y, w1.y);
    } else {         // row-major slot, W(r,c) at [r*LDW + c]: lanes 4 column pairs x 8 row pairs, piece = 16 rows x 8 columns
      cc = (sub & 3) * 8 + 2 * (lane & 3);
      rr = (sub >> 2) * 16 + 2 * (lane >> 2);
      t0 = *reinterpret_cast<const double2*>(&P[rr * LDW + cc]);
      t1 = *reinterpret_cast<const double2*>(&P[(rr + 1) * LDW + cc]);
      w0 = make_double2(t0.x, t1.x);
      w1 = make_double2(t0.y, t1.y);
    }
    const int r = bi * SB + rr, c = bj * SB + cc;
    *reinterpret_cast<double2*>(&Dinv[r + c * NB]) = w0;
    *reinterpret_cast<double2*>(&Dinv[r + (c + 1) * NB]) = w1;
    *reinterpret_cast<double2*>(&DinvT[c + r * NB]) = t0;
    *reinterpret_cast<double2*>(&DinvT[c + (r + 1) * NB]) = t1;
  }
  if (warp == 0) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) logdet += __shfl_xor_sync(FULL, logdet, off);
    if (lane == 0) g.logdet_part[(int64_t)gp * g.J + j] = logdet;
  }
}

// ---------------------------------------------------------------------------------------------
// alpha = L^-T L^-1 ymm by blocked substitution with the inverted diagonal blocks; one CTA per GP.
// ---------------------------------------------------------------------------------------------
constexpr int SOLVE_THREADS = 512;

// sum_{r >= rbeg} Lcol[r] alpha[r] over one column of L by one warp (lanes over rows, two accumulators, ascending rows,
// shuffle tree): the summation order of the backward sweep, shared by k_solve and k_solve_cluster (bit-identical results).
__device__ __forceinline__ double backward_column(const double* Lcol, const double* al, int rbeg, int nv, int lane) {
  double a0 = 0.0, a1 = 0.0;
  int r = rbeg + lane;
  for (; r + 32 < nv; r += 64) {
    a0 = fma(Lcol[r], al[r], a0);
    a1 = fma(Lcol[r + 32], al[r + 32], a1);
  }
  for (; r < nv; r += 32) a0 = fma(Lcol[r], al[r], a0);
  double s = a0 + a1;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  return s;
}

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
  const int nv = g.nv;
  for (int jb = 0; jb < J; ++jb) {
    const double* Lrow = L + (int64_t)jb * NB + row;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const int kend = (jb * NB + row < nv) ? jb * NB : 0;  // padding rows: z = 0
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
      for (; r + 32 < nv; r += 64) {  // (backward_column, written out: the order k_solve_cluster reproduces)
        a0 = fma(Lcol[r], al[r], a0);
        a1 = fma(Lcol[r + 32], al[r + 32], a1);
      }
      for (; r < nv; r += 32) a0 = fma(Lcol[r], al[r], a0);
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


// ---------------------------------------------------------------------------------------------------------
// The same substitution for passes with fewer GPs than SMs (the reference's one-GP-at-a-time call pattern, straggler
// rounds of the batched optimiser, the 50 GPs per GPU of the 8-GPU strong split): one GP = one thread-block CLUSTER
// of C CTAs instead of one CTA.  A single CTA streams the 32 MB factor of an n = 2000 GP at ~25 GB/s (latency bound,
// ~1.1 ms); C CTAs stream it C times as fast.
//   forward   block row i belongs to CTA i mod C; its thread (row, k-partition) keeps the four running sums of
//             k_solve in registers and adds block column j when z_j arrives; the owner of block row j finishes z_j with
//             the inverted diagonal block and writes it into every CTA's copy through distributed shared memory
//   backward  the columns of block column j are dealt over the CTAs; the owner-less diagonal solve is replicated
// Every thread adds its terms in exactly the order of k_solve, so both kernels give bit-identical results (a GP's
// numbers do not depend on how many GPs share its pass - tests/test_gpu_parity.py::test_results_do_not_depend_...).
// ---------------------------------------------------------------------------------------------------------
template <int MAXOWN>  // block rows / columns per CTA: J <= MAXOWN * C
__global__ void __launch_bounds__(SOLVE_THREADS) k_solve_cluster(SolveArgs g, int C) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* z = reinterpret_cast<double*>(smem_raw);  // [npad] full copies, filled block by block by the owners
  double* al = z + g.npad;                            // [npad]
  double* red = al + g.npad;                          // [4][NB]
  double* rv = red + 4 * NB;                          // [NB]
  const int q = (int)cluster.block_rank();
  const int gp = g.list ? g.list[blockIdx.y] : blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t npad = g.npad;
  const int J = g.J, nv = g.nv;
  if (g.fail[gp] != 0) {  // uniform over the cluster: nobody reaches a cluster barrier
    if (q == 0 && tid == 0) g.mll[gp] = -__longlong_as_double(0x7ff0000000000000LL);
    return;
  }
  const double* L = g.Lm + (int64_t)gp * g.mat_stride;
  const double* Dinv = g.Dinv + (int64_t)gp * g.dinv_stride;
  const double* y = g.ymm + (int64_t)gp * npad;
  const int row = tid & (NB - 1), part = tid >> 7;

  // ---- forward
  double fa[MAXOWN][4];
#pragma unroll
  for (int o = 0; o < MAXOWN; ++o) fa[o][0] = fa[o][1] = fa[o][2] = fa[o][3] = 0.0;
  for (int j = 0; j < J; ++j) {
    if (q == j % C) {  // finish z_j (same reduction tree as k_solve) and hand it to every CTA of the cluster
      const int o = j / C;
      double s4 = 0.0;
#pragma unroll
      for (int oo = 0; oo < MAXOWN; ++oo)
        if (oo == o) s4 = (fa[oo][0] + fa[oo][1]) + (fa[oo][2] + fa[oo][3]);
      red[part * NB + row] = s4;
      __syncthreads();
      if (tid < NB) rv[tid] = y[j * NB + tid] - ((red[tid] + red[NB + tid]) + (red[2 * NB + tid] + red[3 * NB + tid]));
      __syncthreads();
      const double* Dj = Dinv + (int64_t)j * NB * NB;
      double s = 0.0;
      for (int kk = part; kk <= row; kk += 4) s = fma(Dj[row + kk * NB], rv[kk], s);
      red[part * NB + row] = s;
      __syncthreads();
      if (tid < NB) {
        const double zv = (red[tid] + red[NB + tid]) + (red[2 * NB + tid] + red[3 * NB + tid]);
        for (int r = 0; r < C; ++r) cluster.map_shared_rank(z, r)[j * NB + tid] = zv;
      }
    }
    cluster.sync();  // z_j visible everywhere (release / acquire at cluster scope)
#pragma unroll
    for (int o = 0; o < MAXOWN; ++o) {
      const int i = q + o * C;
      if (i < J && i > j && i * NB + row < nv) {  // padding rows keep z = 0 (kend = 0 in k_solve)
        const double* Lrow = L + (int64_t)i * NB + row + (int64_t)(j * NB + part) * npad;
        const double* zk = z + j * NB + part;
        double l[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) l[t] = Lrow[(int64_t)(4 * t) * npad];
#pragma unroll
        for (int t = 0; t < 32; ++t) fa[o][t & 3] = fma(l[t], zk[4 * t], fa[o][t & 3]);
      }
    }
  }

  // ---- backward: the 128 columns of block column jb are dealt over the CTAs (column c -> CTA c mod C); a warp sums its
  // column over ALL rows below the block in k_solve's order, the differences z - sum go to every CTA's copy (double
  // buffered: a CTA can be one step ahead of the slowest), and every CTA finishes alpha_jb with the inverted diagonal
  // block for itself - one cluster barrier per block column
  for (int jb = J - 1; jb >= 0; --jb) {
    double* rvb = (jb & 1) ? red : rv;
    const int rbeg = (jb + 1) * NB;
    for (int c = q + C * warp; c < NB; c += C * (SOLVE_THREADS / 32)) {
      const double s = backward_column(L + (int64_t)(jb * NB + c) * npad, al, rbeg, nv, lane);
      const double v = z[jb * NB + c] - s;
      if (lane < C) cluster.map_shared_rank(rvb, lane)[c] = v;  // lane r writes CTA r's copy
    }
    cluster.sync();
    const double* Dj = Dinv + (int64_t)jb * NB * NB;
    for (int c = warp; c < NB; c += SOLVE_THREADS / 32) {
      double s = 0.0;
      for (int kk = c + lane; kk < NB; kk += 32) s = fma(Dj[kk + c * NB], rvb[kk], s);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (lane == 0) al[jb * NB + c] = s;
    }
    __syncthreads();
  }

  if (q != 0) return;
  double* zout = g.zbuf + (int64_t)gp * npad;
  double* aout = g.alpha + (int64_t)gp * npad;
  for (int r = tid; r < (int)npad; r += SOLVE_THREADS) { zout[r] = z[r]; aout[r] = al[r]; }
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

constexpr size_t DIAG_SMEM = (size_t)(NSLOT * SBLK) * sizeof(double);  // 92160 B: two CTAs per SM, or one next to a k_tile_gemm CTA

// per-device opt-in, called by gprb_init for the device of every context (see configure_tile_gemm)
int configure_diag_factor() {
  cudaError_t e = cudaFuncSetAttribute(k_diag_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_diag_factor)", __FILE__, __LINE__);
  return 0;
}

int launch_diag_factor(const DiagArgs& a, int count, cudaStream_t stream) {
  if (count <= 0) return 0;
  const size_t smem = DIAG_SMEM;
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
  // fewer GPs than SMs: one cluster of C CTAs per GP (bit-identical results, see k_solve_cluster)
  const int C = std::min(8, a.J);
  const int maxown = (a.J + C - 1) / C;
  if (count < a.cluster_below && C >= 2 && maxown <= 4) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C, count);
    cfg.blockDim = dim3(SOLVE_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t ec;
#define GPRB_SOLVE_CLUSTER(M)                                                                                       \
    ec = cudaFuncSetAttribute(k_solve_cluster<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    if (ec == cudaSuccess) ec = cudaLaunchKernelEx(&cfg, k_solve_cluster<M>, a, C);
    if (maxown == 1) { GPRB_SOLVE_CLUSTER(1) } else if (maxown == 2) { GPRB_SOLVE_CLUSTER(2) } else { GPRB_SOLVE_CLUSTER(4) }
#undef GPRB_SOLVE_CLUSTER
    if (ec != cudaSuccess) return cuda_fail(ec, "k_solve_cluster launch", __FILE__, __LINE__);
    return 0;
  }
  cudaError_t e = cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_solve)", __FILE__, __LINE__);
  k_solve<<<count, SOLVE_THREADS, smem, stream>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_solve launch", __FILE__, __LINE__);
  return 0;
}

}  // namespace gprb
