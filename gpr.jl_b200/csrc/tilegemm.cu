// Batched fp64 tile GEMM (NT form) on the DMMA tensor pipe - the engine of the blocked
// Cholesky (row a6), the triangular inverse and the LAUUM product (row a8) of SURVEY.md section 8a.
//
// One CTA = one 128x128 output tile of one GP:
//     acc = sum_{k-blocks} Aop[i-rows, k] * Bop[j-rows, k]^T          (both operands column-major, ld = npad)
// Warp-specialised, three warpgroups: warps 0-7 are DMMA consumers with 64x32 register tiles
// (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4, the native fp64 tensor shape on sm_100a); warp 8 is the producer
// (1 KB bulk copies through the TMA engine, one operand column each, completion counted on mbarriers); warps
// 9-11 only exist to complete its warpgroup.  Register reallocation (setmaxnreg): the CTA launches with 168
// registers per thread, the producer warpgroup shrinks to 40 and the two consumer warpgroups grow to 232.  With
// the plain 168-register cap (nine warps = three on one SMSP) the compiler streams the A fragments through a
// single register pair and the LDS latency is exposed to the DMMA pipe (0.90 of its rate in the k-loop); with
// 232 registers the fragments are double-buffered a whole k4-step ahead.
// Padded smem rows (132 doubles) make every fragment load bank-conflict free.
//
// Modes (tile coordinates and k-range derive from `mode`, `step` and blockIdx.x):
//   CHOL_DIAG  S(j,j)   = K(j,j) - sum_{k<j} L(j,k) L(j,k)^T                      -> Lm(j,j)   (potf2 follows)
//   CHOL_COL   L(i,j)   = [K(i,j) - sum_{k<j} L(i,k) L(j,k)^T] * inv(L_jj)^T      -> Lm(i,j)   i > j = step
//   TRTRI_ROW  W(i,j)   = -inv(L_ii) * sum_{k=j}^{i-1} L(i,k) W(k,j)   stored as V(j,i) = W(i,j)^T, i = step
//   LAUUM      Kinv(i,j) = sum_{k>=i} V(i,k) V(j,k)^T  (i >= j)  -> upper tile (j,i) of A (un-transposed) / KinvD(i)
//   CHOL_PANEL L(i,j)   = S(i,j) * inv(L_jj)^T                       in place in Lm, i > j = step      } right-looking
//   CHOL_TRAIL S(i,k)  -= L(i,j) L(k,j)^T   for all j < k <= i          in place in Lm, j = step          } variant
//              (small batches: 3 short launches per block column instead of serial k-loops that grow with the column
//               index - the dependent chain of one factorisation drops from ~4.7 ms to ~1.8 ms at n = 2000)
//   FWD_ROW    T(i,:)   = inv(L_ii) * [T(i,:) - sum_{k<i} L(i,k) T(k,:)]   in place in the right-hand-side block
//              (blocked forward substitution L^-1 K* of the predictive variance, PDMats whiten!), i = step
// where V = L^-T lives in the strictly-upper tiles of Lm and its diagonal blocks in DinvT.
#include "common.cuh"
#include "kernels.h"

namespace gprb {

// Debug timeline (build with -DGPRB_TIMELINE): warp 0 lane 0 of every tile records globaltimer at phase boundaries.
#ifdef GPRB_TIMELINE
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define GPRB_TL(k) do { if (g.tl && threadIdx.x == 0) g.tl[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + (k)] = gtimer(); } while (0)
#else
#define GPRB_TL(k) do { } while (0)
#endif


constexpr int NSTAGE = 4;
constexpr int STAGE_DOUBLES = 2 * KT * LDS_T;           // A + B operand chunk
constexpr int RBUF_DOUBLES = KT * LDS_T;                // one chunk of the post-multiplier
constexpr int NRBUF = 4;                                // ring depth of the post-multiplier chunks
constexpr int N_CONSUMER_WARPS = 8;
constexpr int GEMM_THREADS = (N_CONSUMER_WARPS + 4) * 32;  // two consumer warpgroups + the producer's warpgroup
static_assert(2 * NSTAGE + 2 * NRBUF <= 16, "barrier block holds 16 mbarriers");
static_assert(NSTAGE * STAGE_DOUBLES == NB * LDS_T, "T tile must exactly reuse the stage ring");
constexpr size_t GEMM_SMEM = (size_t)(NSTAGE * STAGE_DOUBLES + NRBUF * RBUF_DOUBLES) * sizeof(double) + 16 * sizeof(uint64_t);

struct TileCoord {
  int i, j;        // output tile (block row, block col)
  int kb0, kb1;    // k-block range [kb0, kb1)
  int a_diag_kb;   // k-block whose A operand comes from DinvT (-1: none)
  int b_diag_kb;   // same for B
  int post;        // 0 none, 1 right-multiply by Dinv[rblk]^T, 2 left-multiply by -Dinv[rblk]
  int rblk;
  bool use_cin;    // T = Cin - acc, else T = acc
};

__device__ __forceinline__ TileCoord tile_coord(int mode, int step, int J, int bx) {
  TileCoord tc;
  tc.a_diag_kb = tc.b_diag_kb = -1;
  tc.post = 0;
  tc.rblk = 0;
  tc.use_cin = false;
  if (mode == GEMM_CHOL_DIAG) {
    tc.i = tc.j = step; tc.kb0 = 0; tc.kb1 = step; tc.use_cin = true;
  } else if (mode == GEMM_CHOL_COL) {
    tc.i = step + 1 + bx; tc.j = step; tc.kb0 = 0; tc.kb1 = step; tc.use_cin = true; tc.post = 1; tc.rblk = step;
  } else if (mode == GEMM_TRTRI_ROW) {
    tc.i = step; tc.j = bx; tc.kb0 = bx; tc.kb1 = step; tc.b_diag_kb = bx; tc.post = 2; tc.rblk = step;
  } else if (mode == GEMM_CHOL_PANEL) {
    tc.i = step + 1 + bx; tc.j = step; tc.kb0 = 0; tc.kb1 = 0; tc.use_cin = true; tc.post = 1; tc.rblk = step;
  } else if (mode == GEMM_CHOL_TRAIL) {  // bx enumerates the lower tiles of the trailing submatrix, row by row
    int ii = (int)((__fsqrt_rn(8.0f * bx + 1.0f) - 1.0f) * 0.5f);
    while ((ii + 1) * (ii + 2) / 2 <= bx) ++ii;
    while (ii * (ii + 1) / 2 > bx) --ii;
    tc.i = step + 1 + ii; tc.j = step + 1 + (bx - ii * (ii + 1) / 2); tc.kb0 = step; tc.kb1 = step + 1; tc.use_cin = true;
  } else if (mode == GEMM_FWD_ROW) {
    tc.i = step; tc.j = bx; tc.kb0 = 0; tc.kb1 = step; tc.use_cin = true; tc.post = 2; tc.rblk = step;
  } else {  // GEMM_LAUUM: bx enumerates (i, j), j <= i, row by row => longest k-range first
    int i = (int)((__fsqrt_rn(8.0f * bx + 1.0f) - 1.0f) * 0.5f);  // approximate, corrected by the two loops below
    while ((i + 1) * (i + 2) / 2 <= bx) ++i;
    while (i * (i + 1) / 2 > bx) --i;
    tc.i = i; tc.j = bx - i * (i + 1) / 2; tc.kb0 = i; tc.kb1 = J; tc.a_diag_kb = i;
    tc.b_diag_kb = (tc.j == i) ? i : -1;
  }
  return tc;
}

enum { SEL_FULL = 0, SEL_SKIP, SEL_MI2, SEL_MI4, SEL_MI6, SEL_NI1, SEL_NI2, SEL_NI3, SEL_TRI0, SEL_TRI4,
       SEL_MLO2, SEL_MLO4, SEL_MLO6, SEL_MLO2_NI3, SEL_MLO4_NI3, SEL_MLO6_NI3, SEL_NLO2 };

// One k-chunk (KT = 16) of a warp's 64x32 register tile restricted, at compile time, to the 8x8 blocks
// (mi, ni) with MI_LO <= mi < MI_LIM, NI_LO <= ni < NI_LIM and mi >= ni + OFF: straight-line unpredicated DMMAs.
template <int MI_LIM, int NI_LIM, int OFF, int MI_LO = 0, int NI_LO = 0>
__device__ __forceinline__ void chunk_mma(double (&acc)[8][4][2], const double* ap, const double* bp) {
  // ap / bp: this lane's fragment pointers at k4 = 0, i.e. base + t * LDS_T + row0 (resp. col0); both operands are
  // k-major in smem with row stride LDS_T
#pragma unroll
  for (int k4 = 0; k4 < KT / 4; ++k4, ap += 4 * LDS_T, bp += 4 * LDS_T) {
    double a[8], b[4];
#pragma unroll
    for (int mi = MI_LO; mi < MI_LIM; ++mi)
      if (mi >= OFF) a[mi] = ap[mi * 8];
#pragma unroll
    for (int ni = NI_LO; ni < NI_LIM; ++ni)
      if (ni + OFF <= 7) b[ni] = bp[ni * 8];
#pragma unroll
    for (int mi = MI_LO; mi < MI_LIM; ++mi)
#pragma unroll
      for (int ni = NI_LO; ni < NI_LIM; ++ni)
        if (mi >= ni + OFF) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
}

// Dispatch one chunk to the specialised body `sel` (warp-uniform).
__device__ __forceinline__ void chunk_dispatch(int sel, double (&acc)[8][4][2], const double* ap, const double* bp) {
  switch (sel) {
    case SEL_SKIP: break;
    case SEL_MI2: chunk_mma<2, 4, -64>(acc, ap, bp); break;
    case SEL_MI4: chunk_mma<4, 4, -64>(acc, ap, bp); break;
    case SEL_MI6: chunk_mma<6, 4, -64>(acc, ap, bp); break;
    case SEL_NI1: chunk_mma<8, 1, -64>(acc, ap, bp); break;
    case SEL_NI2: chunk_mma<8, 2, -64>(acc, ap, bp); break;
    case SEL_NI3: chunk_mma<8, 3, -64>(acc, ap, bp); break;
    case SEL_TRI0: chunk_mma<8, 4, 0>(acc, ap, bp); break;
    case SEL_TRI4: chunk_mma<8, 4, 4>(acc, ap, bp); break;
    case SEL_MLO2: chunk_mma<8, 4, -64, 2>(acc, ap, bp); break;
    case SEL_MLO4: chunk_mma<8, 4, -64, 4>(acc, ap, bp); break;
    case SEL_MLO6: chunk_mma<8, 4, -64, 6>(acc, ap, bp); break;
    case SEL_MLO2_NI3: chunk_mma<8, 3, -64, 2>(acc, ap, bp); break;
    case SEL_MLO4_NI3: chunk_mma<8, 3, -64, 4>(acc, ap, bp); break;
    case SEL_MLO6_NI3: chunk_mma<8, 3, -64, 6>(acc, ap, bp); break;
    case SEL_NLO2: chunk_mma<8, 4, -64, 0, 2>(acc, ap, bp); break;
    default: chunk_mma<8, 4, -64>(acc, ap, bp); break;
  }
}

__device__ __forceinline__ int sel_cols(int ni_lim) {
  return ni_lim <= 0 ? SEL_SKIP : ni_lim == 1 ? SEL_NI1 : ni_lim == 2 ? SEL_NI2 : ni_lim == 3 ? SEL_NI3 : SEL_FULL;
}
__device__ __forceinline__ int sel_rows(int mi_lim) {  // mi_lim is even (n is padded to 16-row chunks)
  return mi_lim <= 0 ? SEL_SKIP : mi_lim == 2 ? SEL_MI2 : mi_lim == 4 ? SEL_MI4 : mi_lim == 6 ? SEL_MI6 : SEL_FULL;
}

// Consumer side of one tile (warps 0-7).  RAGGED = the tile touches the padded tail of the last block row:
// only then are the DMMAs / loads / stores predicated per 8-row slab (mi < mi_valid); full tiles run the clean loop.
template <bool RAGGED>
__device__ __forceinline__ void consume_tile(const GemmArgs& g, const TileCoord& tc, double* stages, double* rbuf,
                                             uint64_t* full, uint64_t* empty, uint64_t* rfull, uint64_t* rempty,
                                             int gp, int nchunks, int rows_valid) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t npad = g.npad;
  // SMSP s = warp & 3 hosts warps {s, s+4}; pair column groups {0,3} / {1,2} on one SMSP so the
  // triangular skips of the post-multiply balance across the four DMMA pipes.
  const int s4 = warp & 3, h = warp >> 2;
  const int wm = h;                    // each SMSP hosts one wm = 0 and one wm = 1 warp (row-padding skip balances)
  const int wn = h ? (3 - s4) : s4;    // ... with column groups {s, 3-s} (triangular post-multiply skip balances)
  const int mi_valid = RAGGED ? min(8, max(0, (rows_valid - wm * 64 + 7) / 8)) : 8;
  const int gq = lane >> 2, t = lane & 3;
  const bool fwd = g.mode == GEMM_FWD_ROW;
  // Column layout of the warp grid.  Default: column group wn owns the four 8-column blocks at 32 wn.  FWD_ROW tiles
  // whose right-hand-side block has fewer than 128 valid test columns (the reference predicts m = 100 test states per
  // step) use a compact layout: the nblk = ceil(ncols / 8) valid 8-column blocks are dealt over the four column groups
  // (extras to groups 0, 1, 3 so the SMSP pairs {0,3} / {1,2} stay balanced) and every warp runs the chunk body
  // specialised to its own block count - the padding columns are neither multiplied nor loaded nor stored.
  // A block may also be split into narrower tiles (g.colw = 64 or 32 columns, one CTA each) when a launch would otherwise
  // have too few tiles to fill the SMs (few, large GPs; the single-GP call pattern): the compact layout then spreads the
  // tile's few column blocks over all eight warps.  Ragged tiles run it with the full-row chunk bodies (the padding rows of
  // the A operand only reach accumulator rows that are never stored or parked).
  int cbase = wn * 32, ni_lim = 4;
  const int ncols_tile = fwd ? min(g.colw, g.ncols - tc.j * g.colw) : NB;
  if (fwd && ncols_tile < NB) {
    const int nblk = (ncols_tile + 7) >> 3, q = nblk >> 2, rem = nblk & 3;
    const int w0 = q + (rem >= 1), w1 = q + (rem >= 2), w2 = q, w3 = q + (rem >= 3);
    ni_lim = wn == 0 ? w0 : wn == 1 ? w1 : wn == 2 ? w2 : w3;
    cbase = 8 * (wn == 0 ? 0 : wn == 1 ? w0 : wn == 2 ? w0 + w1 : w0 + w1 + w2);
  }
  const int row0 = wm * 64 + gq;  // + mi*8
  const int col0 = cbase + gq;    // + ni*8  (operand row index of B)

  // The accumulators start at -Cin (Cholesky tiles: K(i,j); FWD_ROW: the right-hand-side block), loaded straight
  // into the accumulator registers while the first operand chunks are still in flight: the 64 global loads per
  // thread need no extra registers and their latency hides behind the pipeline fill.  After the k-loop
  // acc = sum - Cin = -(Cin - sum); the sign is folded into the stores below.
  const int64_t grow = (int64_t)tc.i * NB, gcol = (int64_t)tc.j * (fwd ? g.colw : NB);
  double acc[8][4][2];
  if (fwd) {
    const double* Tin = g.Tm + (int64_t)(gp - g.t_gp_off) * g.t_stride + gcol + grow * g.ldt;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int r = wm * 64 + mi * 8 + gq, cc = cbase + ni * 8 + 2 * t;
        double2 v = make_double2(0.0, 0.0);
        if ((!RAGGED || mi < mi_valid) && ni < ni_lim) v = *reinterpret_cast<const double2*>(Tin + cc + (int64_t)r * g.ldt);
        acc[mi][ni][0] = -v.x;
        acc[mi][ni][1] = -v.y;
      }
  } else if (tc.use_cin) {
    const double* Cin = g.Cin + (int64_t)gp * g.mat_stride + grow + gcol * npad;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int r = wm * 64 + mi * 8 + gq, cc = wn * 32 + ni * 8 + 2 * t;
        const bool ok = !RAGGED || mi < mi_valid;
        acc[mi][ni][0] = ok ? -Cin[r + (int64_t)cc * npad] : 0.0;
        acc[mi][ni][1] = ok ? -Cin[r + (int64_t)(cc + 1) * npad] : 0.0;
      }
  } else {
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  }

  // Triangular operands: the leading k-block of TRTRI_ROW (B = inv(L_jj)) and of LAUUM (A = inv(L_ii)^T) is
  // triangular, and CHOL_DIAG only needs the lower half of its symmetric tile.  The 8x8 DMMA blocks that would
  // multiply structural zeros (or compute the unused upper half) are skipped through compile-time specialised
  // chunk bodies (no per-DMMA predicates); with the {wn, 3-wn} / {wm 0, 1} warp pairing every SMSP keeps 9/16 of
  // the work of those chunks.  Ragged tiles (last block row) select the row-limited body for their mi_valid slabs.
#ifdef GPRB_TIMELINE
  long long tl_wait = 0;  // cycles this warp spent waiting for operand chunks in the main loop
#endif
  int npred = 0;
  if (g.mode == GEMM_LAUUM) npred = min(nchunks, NB / KT);
  else if (!RAGGED && (g.mode == GEMM_CHOL_DIAG || (g.mode == GEMM_CHOL_TRAIL && tc.i == tc.j))) npred = nchunks;
  else if (!RAGGED && g.mode == GEMM_TRTRI_ROW) npred = NB / KT;
  const int diag_off = 4 * wn - 8 * wm;  // CHOL_DIAG: block (mi, ni) touches the lower triangle iff mi >= ni + diag_off
  // LAUUM diagonal tiles are symmetric: beyond the leading (triangular-operand) k-block only the 8x8 blocks on or below
  // the diagonal are accumulated, the store mirrors them into the upper half (the gradient stage reads whole tiles)
  const bool lsym = !RAGGED && g.mode == GEMM_LAUUM && tc.i == tc.j;
  const int sel_tri = diag_off >= 8 ? SEL_SKIP : diag_off == 4 ? SEL_TRI4 : diag_off == 0 ? SEL_TRI0 : SEL_FULL;
  const int sel_plain = ni_lim < 4 ? sel_cols(ni_lim) : sel_rows(mi_valid);  // column-limited bodies run all 8 row slabs
  {
    int stage = 0; uint32_t phase = 0;
    for (int c = 0; c < nchunks; ++c) {
#ifdef GPRB_TIMELINE
      const long long w0 = clock64();
      mbar_wait(&full[stage], phase);
      if (c > 0) tl_wait += clock64() - w0;  // the wait for the first chunk is the `fill` phase
#else
      mbar_wait(&full[stage], phase);
#endif
      if (c == 0) GPRB_TL(1);
      const double* As = stages + stage * STAGE_DOUBLES;
      int sel = sel_plain;
      if (c < npred) {
        if (g.mode == GEMM_LAUUM) sel = sel_rows(min(mi_valid, 2 * c + 2 - 8 * wm));  // rows r <= 16c + 15 of the triangular A
        else if (g.mode == GEMM_TRTRI_ROW) {                                          // columns <= 16c + 15 of the triangular B
          const int ni_lim = 2 * c + 2 - 4 * wn;
          sel = ni_lim <= 0 ? SEL_SKIP : ni_lim == 2 ? SEL_NI2 : SEL_FULL;
        } else sel = sel_tri;
      } else if (lsym) sel = sel_tri;
      chunk_dispatch(sel, acc, As + t * LDS_T + row0, As + (KT + t) * LDS_T + col0);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
    }
  }

  GPRB_TL(2);
#ifdef GPRB_TIMELINE
  if (g.tl && threadIdx.x == 0) g.tl[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + 7] = (unsigned long long)tl_wait;
#endif
  GPRB_TL(3);
  double* Cout = g.Cout + (int64_t)gp * g.mat_stride;
  if (tc.post == 0) {
    // CHOL_DIAG: S(j,j) -> Lm(j,j).  LAUUM: K^-1(i,j), i > j, goes un-transposed into the free upper tile (j,i) of A
    // (K stays intact in the lower tiles for the gradient stage); diagonal tiles go to the KinvD side buffer.
    double* out = Cout + grow + gcol * npad;
    int64_t ldo = npad;
    const double sgn = tc.use_cin ? -1.0 : 1.0;  // CHOL_DIAG: S = Cin - sum = -acc
    if (g.mode == GEMM_LAUUM) {
      if (tc.i != tc.j) out = Cout + gcol + grow * npad;
      else { out = g.KinvD + (int64_t)gp * g.dinv_stride + (int64_t)tc.i * NB * NB; ldo = NB; }
    }
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        if (RAGGED && mi >= mi_valid) continue;
        if (lsym && mi < ni + diag_off) continue;  // upper 8x8 blocks hold partial sums only: written by their mirror block
        const int r = wm * 64 + mi * 8 + gq, cc = wn * 32 + ni * 8 + 2 * t;
        out[r + (int64_t)cc * ldo] = sgn * acc[mi][ni][0];
        out[r + (int64_t)(cc + 1) * ldo] = sgn * acc[mi][ni][1];
        if (lsym && mi > ni + diag_off)  // strictly-lower block: mirror (rows cc, cc + 1 of column r are adjacent)
          *reinterpret_cast<double2*>(&out[cc + (int64_t)r * ldo]) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
      }
    GPRB_TL(6);
    return;
  }

  // ---- post-multiply: park T in the (now idle) stage ring, then a second DMMA pass against Dinv chunks
  named_bar_sync(1, N_CONSUMER_WARPS * 32);  // every consumer finished reading the ring
  double* Ts = stages;
  if (tc.post == 1) {  // A operand: Ts[k][m] = T[m][k]
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int r = wm * 64 + mi * 8 + gq, cc = wn * 32 + ni * 8 + 2 * t;
        const bool ok = !RAGGED || mi < mi_valid;  // skipped rows are parked as zeros (finite operands for the second pass)
        Ts[cc * LDS_T + r] = ok ? acc[mi][ni][0] : 0.0;
        Ts[(cc + 1) * LDS_T + r] = ok ? acc[mi][ni][1] : 0.0;
      }
  } else {  // B operand: Ts[k][n] = T[k][n]
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int r = wm * 64 + mi * 8 + gq, cc = cbase + ni * 8 + 2 * t;
        const bool ok = !RAGGED || mi < mi_valid;
        if (ni < ni_lim) *reinterpret_cast<double2*>(&Ts[r * LDS_T + cc]) = ok ? make_double2(acc[mi][ni][0], acc[mi][ni][1]) : make_double2(0.0, 0.0);
      }
  }
#pragma unroll
  for (int mi = 0; mi < 8; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  named_bar_sync(1, N_CONSUMER_WARPS * 32);
  GPRB_TL(4);

  // Dinv is lower triangular: R[x][k] == 0 for k > x.  post 1: x = output column, post 2: x = output row.
  const int kmax = (tc.post == 1) ? (wn * 32 + 31) : min(wm * 64 + 63, rows_valid - 1);
  for (int c = 0; c < NB / KT; ++c) {
    const int buf = c % NRBUF;
    mbar_wait(&rfull[buf], (c / NRBUF) & 1);
    if (c * KT <= kmax) {
      const double* Rs = rbuf + buf * RBUF_DOUBLES + t * LDS_T;
      const double* Tc = Ts + (c * KT + t) * LDS_T;
      // triangular skip inside the warp tile: chunk c only meets output columns (post 1) / rows (post 2) x >= 16 c
      int sel = sel_plain;
      if (tc.post == 1) {
        if (sel_plain == SEL_FULL && 2 * c - 4 * wn == 2) sel = SEL_NLO2;
        chunk_dispatch(sel, acc, Tc + row0, Rs + col0);
      } else {
        const int lo = 2 * c - 8 * wm;  // first 8-row slab that meets chunk c: 0, 2, 4 or 6 here (c * KT <= kmax)
        if (lo > 0) {
          if (sel_plain == SEL_FULL) sel = lo == 2 ? SEL_MLO2 : lo == 4 ? SEL_MLO4 : SEL_MLO6;
          else if (sel_plain == SEL_NI3) sel = lo == 2 ? SEL_MLO2_NI3 : lo == 4 ? SEL_MLO4_NI3 : SEL_MLO6_NI3;
        }
        chunk_dispatch(sel, acc, Rs + row0, Tc + col0);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&rempty[buf]);
  }
  GPRB_TL(5);

  if (tc.post == 1) {
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        if (RAGGED && mi >= mi_valid) continue;
        const int r = wm * 64 + mi * 8 + gq, cc = wn * 32 + ni * 8 + 2 * t;
        Cout[grow + r + (gcol + cc) * npad] = -acc[mi][ni][0];      // parked operand was -(Cin - sum)
        Cout[grow + r + (gcol + cc + 1) * npad] = -acc[mi][ni][1];
      }
  } else {  // W(i,j) = -acc, stored transposed as V(j,i); FWD_ROW (parked operand sum - T(i,:)): row r of the solved block
    double* outp = fwd ? g.Tm + (int64_t)(gp - g.t_gp_off) * g.t_stride + gcol + grow * g.ldt : Cout + gcol + grow * npad;
    const int64_t ldo = fwd ? g.ldt : npad;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        if ((RAGGED && mi >= mi_valid) || ni >= ni_lim) continue;
        const int r = wm * 64 + mi * 8 + gq, cc = cbase + ni * 8 + 2 * t;
        *reinterpret_cast<double2*>(&outp[cc + r * ldo]) = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
      }
  }
  GPRB_TL(6);
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) k_tile_gemm(GemmArgs g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stages = reinterpret_cast<double*>(smem_raw);
  double* rbuf = stages + NSTAGE * STAGE_DOUBLES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(rbuf + NRBUF * RBUF_DOUBLES);
  uint64_t* full = bars;            // [NSTAGE]
  uint64_t* empty = bars + NSTAGE;  // [NSTAGE]
  uint64_t* rfull = bars + 2 * NSTAGE;            // [NRBUF]
  uint64_t* rempty = bars + 2 * NSTAGE + NRBUF;   // [NRBUF]

  GPRB_TL(0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gp = g.list ? g.list[blockIdx.y] : g.gp_off + blockIdx.y;
  // A GP whose factorisation already hit a non-positive pivot (make_posdef! will retry it with more jitter) skips the
  // rest of the failed attempt, like dpotrf stopping at the failing column.  The load overlaps the setup below.
  const int failed = g.fail ? g.fail[gp] : 0;
  const TileCoord tc = tile_coord(g.mode, g.step, g.J, blockIdx.x);
  const int64_t npad = g.npad;
  const double* Lm = g.Lm + (int64_t)gp * g.mat_stride;
  const double* DinvT = g.DinvT + (int64_t)gp * g.dinv_stride;
  const double* Dinv = g.Dinv + (int64_t)gp * g.dinv_stride;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], N_CONSUMER_WARPS); }
    for (int s = 0; s < NRBUF; ++s) { mbar_init(&rfull[s], 1); mbar_init(&rempty[s], N_CONSUMER_WARPS); }
    mbar_fence_init();
  }
  if (failed != 0) return;  // uniform over the CTA
  __syncthreads();

  // Padding skip: rows/cols >= nv (n rounded up to the 16-wide chunk) are never computed and never read.
  // Only the last block row / the last k-block are ragged.
  const int nvl = g.nv - (g.J - 1) * NB;                          // valid rows of the last 128-block (multiple of 16)
  const int last_kb_chunks = nvl / KT;                             // chunks of k-block J-1 that hold valid k
  const int nchunks = (tc.kb1 - tc.kb0) * (NB / KT) - ((tc.kb1 == g.J && tc.kb1 > tc.kb0) ? (NB / KT - last_kb_chunks) : 0);
  const int rows_valid = (tc.i == g.J - 1) ? nvl : NB;            // valid output rows of this tile

  if (warp >= N_CONSUMER_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");  // whole producer warpgroup; frees registers for the consumers
    if (warp != N_CONSUMER_WARPS) return;
    // ===================== producer warp =====================
    int stage = 0; uint32_t phase = 0;
    int rissued = 0;
    auto issue_r = [&](int c) {  // chunk c of the post-multiplier -> rbuf[c & 1]
      const int buf = c % NRBUF;
      mbar_wait(&rempty[buf], ((c / NRBUF) & 1) ^ 1);
      // one lane issues all the copies of a chunk with warp-uniform operands: per-lane addresses would make the
      // compiler serialise the uniform-datapath UBLKCP through an ELECT / R2UR.BROADCAST loop (~90 cycles per copy)
      if (lane == 0) {
        mbar_expect_tx(&rfull[buf], KT * NB * sizeof(double));
        const double* src = Dinv + (int64_t)tc.rblk * NB * NB + (int64_t)(c * KT) * NB;
        double* dst = rbuf + buf * RBUF_DOUBLES;
#pragma unroll
        for (int r = 0; r < KT; ++r) bulk_g2s(dst + r * LDS_T, src + r * NB, NB * sizeof(double), &rfull[buf]);
      }
    };
    int issued = 0;  // main chunks issued so far; the post-multiplier prefetch follows the first ring fill
    // bytes of one B-operand row: a narrow FWD_ROW tile only owns colw columns of the right-hand-side row
    const uint32_t bbytes = (uint32_t)((g.mode == GEMM_FWD_ROW ? g.colw : NB) * sizeof(double));
    for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
      const double* srcA; const double* srcB; int64_t ldA, ldB;
      if (kb == tc.a_diag_kb) { srcA = DinvT + (int64_t)kb * NB * NB; ldA = NB; }
      else { srcA = Lm + (int64_t)tc.i * NB + (int64_t)kb * NB * npad; ldA = npad; }
      if (kb == tc.b_diag_kb) { srcB = DinvT + (int64_t)kb * NB * NB; ldB = NB; }
      else if (g.mode == GEMM_FWD_ROW) { srcB = g.Tm + (int64_t)(gp - g.t_gp_off) * g.t_stride + (int64_t)tc.j * g.colw + (int64_t)kb * NB * g.ldt; ldB = g.ldt; }
      else { srcB = Lm + (int64_t)tc.j * NB + (int64_t)kb * NB * npad; ldB = npad; }
      const int cend = (kb == g.J - 1) ? last_kb_chunks : NB / KT;
      for (int c = 0; c < cend; ++c) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (lane == 0) {
          mbar_expect_tx(&full[stage], KT * (NB * sizeof(double) + bbytes));
          double* dst = stages + stage * STAGE_DOUBLES;
          const double* sa = srcA + (int64_t)(c * KT) * ldA;
          const double* sb = srcB + (int64_t)(c * KT) * ldB;
#pragma unroll
          for (int r = 0; r < KT; ++r) {
            bulk_g2s(dst + r * LDS_T, sa + r * ldA, NB * sizeof(double), &full[stage]);
            bulk_g2s(dst + (KT + r) * LDS_T, sb + r * ldB, bbytes, &full[stage]);
          }
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        if (++issued == NSTAGE && tc.post) { for (; rissued < NRBUF; ++rissued) issue_r(rissued); }
      }
    }
    if (tc.post) for (; rissued < NB / KT; ++rissued) issue_r(rissued);
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");  // consumer warpgroups
  if (rows_valid == NB) consume_tile<false>(g, tc, stages, rbuf, full, empty, rfull, rempty, gp, nchunks, rows_valid);
  else consume_tile<true>(g, tc, stages, rbuf, full, empty, rfull, rempty, gp, nchunks, rows_valid);
}

// The dynamic shared-memory opt-in is a per-device function attribute: gprb_init calls this for the device of every
// context it creates (a process-wide "configured" flag would leave a second GPU of the same process un-opted-in).
int configure_tile_gemm() {
  cudaError_t e = cudaFuncSetAttribute(k_tile_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_tile_gemm)", __FILE__, __LINE__);
  return 0;
}

int launch_tile_gemm(const GemmArgs& g, int ntiles, int count, cudaStream_t stream) {
  if (ntiles <= 0 || count <= 0) return 0;
  dim3 grid(ntiles, count);
  k_tile_gemm<<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(g);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_tile_gemm launch", __FILE__, __LINE__);
  return 0;
}

}  // namespace gprb
