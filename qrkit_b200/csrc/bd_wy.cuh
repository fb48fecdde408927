// bd_wy.cuh — blocked (compact-WY) FP64 Householder QR of medium diagonal blocks, one CTA per block, the
// trailing update on the FP64 tensor cores (mma.sync m8n8k4 f64 = SASS DMMA.8x8x4).
//
// Reference path: BlockDiagonalSparseQR::factorize for dynamically sized blocks
// (src/QRKit/BlockDiagonalSparseQR.h:432-526 with BlockQRSolver = HouseholderQR<MatrixXd>), and the panel /
// compact-WY machinery of BlockedThinQRBase (src/QRKit/BlockedThinQRBase.h:308-333: Y, T and the trailing update
// A[:,j] += Y (T^T (Y^T A[:,j])), which the reference performs one column at a time as GEMVs).  Here the same
// update is three small GEMMs per 8-column tile, A2^T -= ((A2^T V) T) V^T, issued as DMMA.  BASELINE config 5.
//
// Layout.  The block lives in shared memory column-major with leading dimension ld = rp + 4 (rp = rows padded
// to a multiple of 8, so ld = 4 mod 8): the two access patterns of the DMMA fragments,
//   (row + q, col + g) and (row + g', col + q)   with q = lane & 3, g = lane >> 2,
// then hit 16 distinct 8-byte banks per half-warp.  Inside an 8-row tile the K/N slots of the fragments are
// permuted (kappa below) so that the SAME two registers of a thread are the A operand of the first GEMM and the
// accumulator of the last one: a tile of A2 is read once and written once per panel, and V is read in place.
//
// Schedule.  Panels of 8 columns.  One warp factors a panel in registers (rows across lanes, ONE batched
// warp-shuffle all-reduce per column: tail dot products of the pivot column with itself and with the columns to
// its right), writes the packed columns back, forms S = V^T V with DMMA and the 8x8 T by the dlarft recurrence
// (one lane per row of T).  In phase p every warp applies panel p to its 8-column tiles; the owner of the next
// panel takes tile p+1 first and factors it at once (look-ahead) while the others finish the phase.
// The right-hand side rides along as one more (virtual) tile, then R x = (Q^T b)[0:c] is solved by warp 0.
#pragma once
#include "bd_generic.cuh"

namespace qrk {

// Optional cycle trace of CTA 0 (tools/wy_trace.cu compiles with -DQRK_WY_TRACE); compiled out of the library.
#ifdef QRK_WY_TRACE
__device__ long long g_wy_trace[4096];
__device__ int g_wy_trace_n;
#define WY_TRACE(tag)                                                                     \
  do {                                                                                    \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {                                     \
      const int slot = atomicAdd(&g_wy_trace_n, 1);                                       \
      if (slot < 2048) { g_wy_trace[2 * slot] = (tag) * 16 + (threadIdx.x >> 5); g_wy_trace[2 * slot + 1] = clock64(); } \
    }                                                                                     \
  } while (0)
#define WY_CLK(i) do { const long long now_ = clock64(); wy_acc[i] += now_ - wy_last; wy_last = now_; } while (0)
#define WY_TRACE_VAL(tag, val)                                                            \
  do {                                                                                    \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {                                     \
      const int slot = atomicAdd(&g_wy_trace_n, 1);                                       \
      if (slot < 2048) { g_wy_trace[2 * slot] = (tag) * 16 + (threadIdx.x >> 5); g_wy_trace[2 * slot + 1] = (val); } \
    }                                                                                     \
  } while (0)
#else
#define WY_TRACE(tag) do { } while (0)
#define WY_CLK(i) do { } while (0)
#endif

__device__ unsigned g_wy_sm_ticket[256];   // see bd_wy_factor_kernel: spreads panel warps over the SM sub-partitions

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct WyGeom {
  int rp, cp, ld;
  __host__ __device__ WyGeom(int r, int c) {
    cp = (c + 7) & ~7;
    const int r8 = (r + 7) & ~7;
    rp = r8 > cp ? r8 : cp;
    ld = rp + 4;
  }
};
// panel warp scratch: two step buffers of the column-per-lane panel (raw pivot column 4 x (RQ + 2), pivot row 8, norm partials 4;
// RQ = 32 rows per lane at most) and the 8x8 S = V^T V; the cooperative look-ahead apply reuses it for its partial W (W x 64)
constexpr int kWyStepBuf = 4 * (32 + 2) + 12;
constexpr int kWyScratch = 2 * kWyStepBuf + 64 + 40;   // (the team-panel experiment needs 396)
constexpr int kWyPB = 96;   // diagonal tile of V (unit lower triangular 8x8), column stride 12 (= 12 mod 16)

__host__ __device__ inline size_t wy_smem_bytes(int r, int c) {
  const WyGeom g(r, c);
  const size_t d = (size_t)g.cp * g.ld + g.ld + 2 * kWyPB + 2 * 64 + kWyScratch + 2 * (size_t)g.cp;
  return d * 8 + 16;
}
// panel rows per lane in units of 8 (template parameter MR; the panel warp holds 8 MR rows of ONE column per lane, the
// apply warps 4 MR row tiles): 1..4; 0 = block too tall for this kernel
__host__ __device__ inline int wy_mr(int r, int c) {
  const WyGeom g(r, c);
  return g.rp <= 32 ? 1 : g.rp <= 64 ? 2 : g.rp <= 96 ? 3 : g.rp <= 128 ? 4 : 0;
}
// Team size.  The panel chain is serial (one warp), the other warps only apply finished panels, and the kernel is
// latency bound: what counts is how many blocks are resident per SM.  Small blocks are limited by registers x threads,
// so they get one or two warps; only blocks whose shared-memory footprint already caps residency get four.
__host__ __device__ inline int wy_warps(int r, int c) {
  const int cp = WyGeom(r, c).cp;
  return cp <= 24 ? 1 : cp <= 48 ? 2 : 4;
}

__device__ __forceinline__ bool lane_is(int j, int q, int jj, int qq) { return j == jj && q == qq; }

// fragment slot -> row inside an 8-row tile: k-step h, slot q  (see header)
__device__ __forceinline__ int wy_kappa(int q, int h) { return h ? 4 + ((q + 2) & 3) : q; }

// Reduce-scatter of N (<= 8) doubles per lane over the warp by recursive halving: each round a lane keeps one half
// of its values and adds the partner's copy of that half; for N = 8 every lane ends with the total of value
// (lane >> 2), at 9 64-bit shuffles and 9 additions instead of the 40 + 40 of a butterfly per value.
template <int N, int OFF>
__device__ __forceinline__ double wy_reduce_scatter(double (&v)[N], int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  if constexpr (N == 1) {
    double x = v[0];
#pragma unroll
    for (int o = OFF; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
  } else {
    constexpr int H = (N + 1) / 2;
    const bool upper = (lane & OFF) != 0;
    double w[H];
#pragma unroll
    for (int i = 0; i < H; i++) {
      const double lo = v[i];
      const double hi = (H + i < N) ? v[(H + i < N) ? H + i : 0] : 0.0;
      const double recv = __shfl_xor_sync(FULL, upper ? lo : hi, OFF);
      w[i] = (upper ? hi : lo) + recv;
    }
    return wy_reduce_scatter<H, OFF / 2>(w, lane);
  }
}
// ---- panel factorisation: columns p..p+7, rows p..rp-1, by ONE warp -------------------------------------------
// One column step; K is a template parameter so that every register-array index is a compile-time constant.
//
// The step needs ONE warp all-reduce: d_j = sum_{rows > K} a_K a_j for ALL eight columns of the panel.  For j >= K
// these are Eigen's tail quantities (d_K = tailSqNorm, d_j -> v^T a_j); for j < K column j already holds the
// essential part of v_j, so d_j is the cross term of S = V^T V that the T factor needs: T comes out of the same
// reduction and is finished when the last column is (no separate V^T V pass).  The sums are reduce-scattered with
// shuffles (9 instead of 40), the eight totals and the pivot row travel through 32 doubles of shared memory.
// Scalar chain (all lanes redundantly): 1/||x|| by MUFU.RSQ64H + 2 Newton steps; the reciprocal of x0 - beta is
// seeded from the half-converged norm so that MUFU.RCP64H overlaps the second Newton step; tau v^T a_j is formed as
// -(d a_Kj + t_j) / beta so that it does not wait for that reciprocal.
template <int MR, int K>
__device__ __forceinline__ void wy_panel_step(double (&a)[MR][8], double* tau_out, double (&Trow)[8], double* scratch,
                                              int lane
#ifdef QRK_WY_TRACE
                                              , long long (&wy_acc)[6], long long& wy_last
#endif
                                              ) {
  double* gbuf = scratch + (K & 1) * 16;      // [0..8): reduced sums, [8..16): pivot row
  double part[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < MR; m++) {
      const double ak = (m > 0 || lane > K) ? a[m][K] : 0.0;
      s = fma(ak, a[m][j], s);
    }
    part[j] = s;
  }
  WY_CLK(0);
  const double mine = wy_reduce_scatter<8, 16>(part, lane);
  WY_CLK(1);
  if ((lane & 3) == 0) gbuf[lane >> 2] = mine;
  __syncwarp();
  double d[8], row[8];
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    const double2 u = *reinterpret_cast<const double2*>(gbuf + j);
    const double2 w = *reinterpret_cast<const double2*>(gbuf + 8 + j);
    d[j] = u.x; d[j + 1] = u.y; row[j] = w.x; row[j + 1] = w.y;
  }
  WY_CLK(2);
  // Eigen makeHouseholder: beta = -sign(x0) ||x||, ess = tail / (x0 - beta), tau = (beta - x0) / beta
  const double c0 = row[K];
  const bool degenerate = d[K] <= DBL_MIN;
  const double nsq = fma(c0, c0, d[K]);
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(nsq));
  const double hx = 0.5 * nsq, ac0 = fabs(c0);
  // (depth-minimised as householder_scalars, common.cuh: norm from ONE Newton step + the Heron correction, the second Newton
  //  step only feeds 1/beta beside the chain; reciprocal seeded from the 20-bit norm and finished by one cubic step)
  double r0;
  { const double n0 = fma(nsq, y0, ac0); asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(n0)); }   // seed only
  double y1;
  { const double t = y0 * y0; const double e = fma(-hx, t, 0.5); y1 = fma(y0, e, y0); }
  double norm = nsq * y1;
  norm = fma(fma(-norm, norm, nsq), 0.5 * y1, norm);
  double y;
  { const double t = y1 * y1; const double e = fma(-hx, t, 0.5); y = fma(y1, e, y1); }
  const double dabs = ac0 + norm;              // |x0 - beta|
  double rabs;
  { const double e = fma(-dabs, r0, 1.0); rabs = fma(r0, fma(e, e, e), r0); }
  const bool pos = c0 >= 0.0;
  double beta = pos ? -norm : norm;
  double ib = pos ? -y : y;                    // 1 / beta
  double dd = pos ? dabs : -dabs;              // x0 - beta
  double inv = pos ? rabs : -rabs;             // 1 / (x0 - beta)
  double tau = (beta - c0) * ib;
  if (degenerate) { inv = 0.0; tau = 0.0; beta = c0; ib = 0.0; dd = 0.0; }
  WY_CLK(3);
  // rows >= K of the columns to the right: a_j -= (tau v^T a_j) v, with v = [1; tail * inv] = mult * inv
  double mult[MR];
#pragma unroll
  for (int m = 0; m < MR; m++) mult[m] = (m > 0 || lane > K) ? a[m][K] : (lane == K) ? dd : 0.0;
#pragma unroll
  for (int j = K + 1; j < 8; j++) {
    const double sj = -fma(dd, row[j], d[j]) * ib;     // tau v^T a_j
    const double sinv = sj * inv;
#pragma unroll
    for (int m = 0; m < MR; m++) a[m][j] = fma(-mult[m], sinv, a[m][j]);
  }
  // column K of T (lane i < 8 owns row i): S_iK = v_i[K] + inv d_i, T_iK = -tau sum_m T_im S_mK, T_KK = tau
  {
    double sum = 0.0;
#pragma unroll
    for (int m = 0; m < K; m++) sum = fma(Trow[m], fma(inv, d[m], row[m]), sum);
    Trow[K] = (lane == K) ? tau : (lane < K) ? -tau * sum : 0.0;
  }
  // column K: beta on the diagonal, essential part below
#pragma unroll
  for (int m = 0; m < MR; m++) {
    if (m == 0) a[0][K] = (lane > K) ? a[0][K] * inv : (lane == K) ? beta : a[0][K];
    else a[m][K] *= inv;
  }
  if (lane == K) tau_out[K] = tau;
  if (K < 7) {
    if (lane == K + 1) {
      double* nb = scratch + ((K + 1) & 1) * 16 + 8;
#pragma unroll
      for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(nb + j) = make_double2(a[0][j], a[0][j + 1]);
    }
  }
  WY_CLK(4);
}

template <int MR>
__device__ __forceinline__ void wy_factor_panel(double* sA, int ld, int rp, int p, double* PB0, double* sT, double* sS,
                                                double* sTau, int lane) {
  const int nrow = rp - p;
  WY_TRACE(0);
  double a[MR][8];
  {
    const double* base = sA + p * ld + p + lane;
#pragma unroll
    for (int m = 0; m < MR; m++) {
      const bool ok = lane + 32 * m < nrow;
#pragma unroll
      for (int j = 0; j < 8; j++) a[m][j] = ok ? base[j * ld + 32 * m] : 0.0;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(sS + 8 + j) = make_double2(a[0][j], a[0][j + 1]);
  }
  double Trow[8];
  double* tauv = sTau + p;
#ifdef QRK_WY_TRACE
  long long wy_acc[6] = {0, 0, 0, 0, 0, 0};
  long long wy_last = clock64();
#define WY_STEP(k) wy_panel_step<MR, k>(a, tauv, Trow, sS, lane, wy_acc, wy_last)
#else
#define WY_STEP(k) wy_panel_step<MR, k>(a, tauv, Trow, sS, lane)
#endif
  WY_TRACE(1);
  WY_STEP(0);
  WY_STEP(1);
  WY_STEP(2);
  WY_STEP(3);
  WY_STEP(4);
  WY_STEP(5);
  WY_STEP(6);
  WY_STEP(7);
  WY_TRACE(2);
#ifdef QRK_WY_TRACE
  for (int i = 0; i < 5; i++) WY_TRACE_VAL(12 + i, wy_acc[i]);
#endif
  // packed columns back in place (R above / on the diagonal, essential parts below), unit-lower diagonal tile of V, T
  {
    double* base = sA + p * ld + p + lane;
#pragma unroll
    for (int m = 0; m < MR; m++) {
      if (lane + 32 * m < nrow) {
#pragma unroll
        for (int j = 0; j < 8; j++) base[j * ld + 32 * m] = a[m][j];
      }
    }
  }
  if (lane < 8) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      PB0[j * 12 + lane] = (lane < j) ? 0.0 : (lane == j) ? 1.0 : a[0][j];
      sT[lane * 8 + j] = Trow[j];
    }
  }
  __syncwarp();
  WY_TRACE(5);
}

// ---------------------------------------------------------------------------------------------------------------
// Gram-downdated panel (experimental, -DQRK_WY_GRAM).  The per-column all-reduce above is the longest link of the serial chain.  Here the
// warp reduces ONCE per panel: G = P^T P over the panel rows (36 sums).  Reflections are orthogonal on rows >= k, so the
// Gram of the current columns over rows >= k+1 is G_k+1 = G_k - R_k R_k^T with R_k the finished row k: every later
// tail quantity follows from scalars,  t_j = G_k[k][j] - a_kk a_kj,  tailSq = G_k[k][k] - a_kk^2,
// and a column step shrinks to [row k through shared memory -> scalar chain -> lane-local rank-1 update].
// Accuracy: t_j carries an absolute error eps ||a_k|| ||a_j||, the same order as the rounding of v^T a_j itself;
// tailSq is a difference and loses digits when the tail is short against the head, so whenever tailSq <= 2^-10 G_kk
// (or is not positive) it is re-summed directly from the registers (one butterfly): the norm, hence beta, tau and the
// essential part, keep full relative accuracy.  The Gram lives distributed over the lanes (two entries each).
// S = V^T V for the T factor is one more batched reduction after the last column.
// ---------------------------------------------------------------------------------------------------------------
struct WyRefl { double beta, tau, inv, dd, ib; };

// Eigen makeHouseholder from x0 and tailSqNorm: beta = -sign(x0) ||x||, ess = tail / (x0 - beta), tau = (beta - x0) / beta
__device__ __forceinline__ WyRefl wy_reflector(double c0, double tailSq) {
  const bool degenerate = tailSq <= DBL_MIN;
  const double nsq = fma(c0, c0, tailSq);
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(nsq));
  const double hx = 0.5 * nsq, ac0 = fabs(c0);
  // (depth-minimised as householder_scalars, common.cuh: norm from ONE Newton step + the Heron correction, the second Newton
  //  step only feeds 1/beta beside the chain; reciprocal seeded from the 20-bit norm and finished by one cubic step)
  double r0;
  { const double n0 = fma(nsq, y0, ac0); asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(n0)); }   // seed only
  double y1;
  { const double t = y0 * y0; const double e = fma(-hx, t, 0.5); y1 = fma(y0, e, y0); }
  double norm = nsq * y1;
  norm = fma(fma(-norm, norm, nsq), 0.5 * y1, norm);
  double y;
  { const double t = y1 * y1; const double e = fma(-hx, t, 0.5); y = fma(y1, e, y1); }
  const double dabs = ac0 + norm;              // |x0 - beta|
  double rabs;
  { const double e = fma(-dabs, r0, 1.0); rabs = fma(r0, fma(e, e, e), r0); }
  const bool pos = c0 >= 0.0;
  WyRefl h;
  h.beta = pos ? -norm : norm;
  h.ib = pos ? -y : y;                         // 1 / beta
  h.dd = pos ? dabs : -dabs;                   // x0 - beta
  h.inv = pos ? rabs : -rabs;                  // 1 / (x0 - beta)
  h.tau = (h.beta - c0) * h.ib;
  if (degenerate) { h.inv = 0.0; h.tau = 0.0; h.beta = c0; h.ib = 0.0; h.dd = 0.0; }
  return h;
}

// slot whose total a lane holds after wy_reduce_scatter<n, 16> (-1: padding)
__device__ __forceinline__ int wy_slot_of_lane(int n, int lane) {
  int base = 0;
  for (int off = 16; off > 0 && n > 1; off >>= 1) {
    const int h = (n + 1) >> 1;
    if (lane & off) { base += h; n -= h; } else n = h;
  }
  return n >= 1 ? base : -1;
}
// strictly-upper pair (i < j) number e = j(j-1)/2 + i  ->  (i, j)
__device__ __forceinline__ void wy_pair(int e, int& i, int& j) {
  j = 1;
  while ((j * (j + 1)) / 2 <= e) j++;
  i = e - (j * (j - 1)) / 2;
}

template <int MR, int K>
__device__ __forceinline__ void wy_gram_step(double (&a)[MR][8], double* tau_out, double& g0, double& g1, double* scratch, int lane) {
  const double* buf = scratch + (K & 1) * 16;   // [0..8): row K of the downdated Gram, [8..16): panel row K
  __syncwarp();
  double gd[8], row[8];
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    const double2 u = *reinterpret_cast<const double2*>(buf + j);
    const double2 w = *reinterpret_cast<const double2*>(buf + 8 + j);
    gd[j] = u.x; gd[j + 1] = u.y; row[j] = w.x; row[j + 1] = w.y;
  }
  const int i0 = lane >> 3, i1 = 4 + (lane >> 3), j0 = lane & 7;      // this lane owns G[i0][j0] and G[i1][j0]
  const double gx0 = buf[i0], gx1 = buf[i1], gxj = buf[j0], rx0 = buf[8 + i0], rx1 = buf[8 + i1], rxj = buf[8 + j0];
  const double c0 = row[K], gkk = gd[K];
  double tailSq = fma(-c0, c0, gkk);
  if (!(tailSq > 0x1p-10 * gkk)) {             // warp-uniform: cancellation, a (numerically) zero column, or no rows left
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < MR; m++) { const double ak = (m > 0 || lane > K) ? a[m][K] : 0.0; s = fma(ak, ak, s); }
    tailSq = warp_sum(s);
  }
  const WyRefl h = wy_reflector(c0, tailSq);
  double mult[MR];
#pragma unroll
  for (int m = 0; m < MR; m++) mult[m] = (m > 0 || lane > K) ? a[m][K] : (lane == K) ? h.dd : 0.0;
#pragma unroll
  for (int j = K + 1; j < 8; j++) {
    const double tj = fma(-c0, row[j], gd[j]);               // sum_{rows > K} a_K a_j
    const double sj = -fma(h.dd, row[j], tj) * h.ib;         // tau v^T a_j
    const double sinv = sj * h.inv;
#pragma unroll
    for (int m = 0; m < MR; m++) a[m][j] = fma(-mult[m], sinv, a[m][j]);
  }
#pragma unroll
  for (int m = 0; m < MR; m++) {
    if (m == 0) a[0][K] = (lane > K) ? a[0][K] * h.inv : (lane == K) ? h.beta : a[0][K];
    else a[m][K] *= h.inv;
  }
  if (lane == K) tau_out[K] = h.tau;
  // downdate the entries this lane owns with the finished row: R_Kx = a_Kx - tau v^T a_x
  auto rk = [&](double rx, double gx) { const double tx = fma(-c0, rx, gx); return fma(fma(h.dd, rx, tx), h.ib, rx); };
  const double Rj = rk(rxj, gxj);
  g0 = fma(-rk(rx0, gx0), Rj, g0);
  g1 = fma(-rk(rx1, gx1), Rj, g1);
  if (K < 7) {
    double* nb = scratch + ((K + 1) & 1) * 16;
    if (K + 1 < 4) { if (i0 == K + 1) nb[j0] = g0; }
    else { if (i1 == K + 1) nb[j0] = g1; }
    if (lane == K + 1) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(nb + 8 + j) = make_double2(a[0][j], a[0][j + 1]);
    }
  }
}

template <int MR>
__device__ __forceinline__ void wy_factor_panel_gram(double* sA, int ld, int rp, int p, double* PB0, double* sT, double* scratch,
                                                     double* sTau, int lane) {
  const int nrow = rp - p;
  WY_TRACE(0);
  double a[MR][8];
  {
    const double* base = sA + p * ld + p + lane;
#pragma unroll
    for (int m = 0; m < MR; m++) {
      const bool ok = lane + 32 * m < nrow;
#pragma unroll
      for (int j = 0; j < 8; j++) a[m][j] = ok ? base[j * ld + 32 * m] : 0.0;
    }
  }
  double* Gfull = scratch + 32;       // 8x8, symmetric
  double* sS = scratch + 96;          // 8x8, strictly upper part used
  // ---- G = P^T P: 28 off-diagonal + 4 diagonal sums in one 32-wide reduce-scatter (lane e ends with entry e), 4 more diagonals
  double g0, g1;
  {
    double v32[32], v4[4];
#pragma unroll
    for (int j = 0; j < 8; j++) {
#pragma unroll
      for (int i = 0; i <= j; i++) {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < MR; m++) s = fma(a[m][i], a[m][j], s);
        if (i < j) v32[(j * (j - 1)) / 2 + i] = s;
        else if (j < 4) v32[28 + j] = s;
        else v4[j - 4] = s;
      }
    }
    const double t32 = wy_reduce_scatter<32, 16>(v32, lane);
    const double t4 = wy_reduce_scatter<4, 16>(v4, lane);
    __syncwarp();                                  // the scratch of an earlier panel is no longer read
    if (lane < 28) { int i, j; wy_pair(lane, i, j); Gfull[i * 8 + j] = t32; Gfull[j * 8 + i] = t32; }
    else Gfull[(lane - 28) * 9] = t32;
    if ((lane & 7) == 0) Gfull[(4 + (lane >> 3)) * 9] = t4;
    __syncwarp();
    const int i0 = lane >> 3, j0 = lane & 7;
    g0 = Gfull[i0 * 8 + j0];
    g1 = Gfull[(4 + i0) * 8 + j0];
    if (i0 == 0) scratch[j0] = g0;                 // step 0 reads row 0 of G and panel row 0
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(scratch + 8 + j) = make_double2(a[0][j], a[0][j + 1]);
    }
  }
  double* tauv = sTau + p;
  WY_TRACE(1);
  wy_gram_step<MR, 0>(a, tauv, g0, g1, scratch, lane);
  wy_gram_step<MR, 1>(a, tauv, g0, g1, scratch, lane);
  wy_gram_step<MR, 2>(a, tauv, g0, g1, scratch, lane);
  wy_gram_step<MR, 3>(a, tauv, g0, g1, scratch, lane);
  wy_gram_step<MR, 4>(a, tauv, g0, g1, scratch, lane);
  wy_gram_step<MR, 5>(a, tauv, g0, g1, scratch, lane);
  wy_gram_step<MR, 6>(a, tauv, g0, g1, scratch, lane);
  wy_gram_step<MR, 7>(a, tauv, g0, g1, scratch, lane);
  WY_TRACE(2);
  // packed columns back in place, unit-lower diagonal tile of V
  {
    double* base = sA + p * ld + p + lane;
#pragma unroll
    for (int m = 0; m < MR; m++) {
      if (lane + 32 * m < nrow) {
#pragma unroll
        for (int j = 0; j < 8; j++) base[j * ld + 32 * m] = a[m][j];
      }
    }
  }
  if (lane < 8) {
#pragma unroll
    for (int j = 0; j < 8; j++) PB0[j * 12 + lane] = (lane < j) ? 0.0 : (lane == j) ? 1.0 : a[0][j];
  }
  WY_TRACE(3);
  // ---- S = V^T V (strictly upper part): one 28-wide reduce-scatter over the rows in registers
  {
    double sv[28];
#pragma unroll
    for (int j = 1; j < 8; j++) {
#pragma unroll
      for (int i = 0; i < j; i++) {
        double s = (lane < j) ? 0.0 : (lane == j) ? a[0][i] : a[0][i] * a[0][j];     // rows 0..31: v_j = [0; 1; essential]
#pragma unroll
        for (int m = 1; m < MR; m++) s = fma(a[m][i], a[m][j], s);
        sv[(j * (j - 1)) / 2 + i] = s;
      }
    }
    const double ts = wy_reduce_scatter<28, 16>(sv, lane);
    const int slot = wy_slot_of_lane(28, lane);
    if (slot >= 0) { int i, j; wy_pair(slot, i, j); sS[i * 8 + j] = ts; }
  }
  __syncwarp();
  WY_TRACE(4);
  // T (upper triangular, Q = I - V T V^T): T_kk = tau_k, T[0:k,k] = -tau_k T[0:k,0:k] S[0:k,k]   (LAPACK dlarft /
  // Eigen make_block_householder_triangular_factor); lane i owns row i
  if (lane < 8) {
    double Tr[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      double sum = 0.0;
#pragma unroll
      for (int m = 0; m < k; m++) sum = fma(Tr[m], sS[m * 8 + k], sum);
      const double tk = tauv[k];
      Tr[k] = (k < lane) ? 0.0 : (k == lane) ? tk : -tk * sum;
      sT[lane * 8 + k] = Tr[k];
    }
  }
  __syncwarp();
  WY_TRACE(5);
}


// ---------------------------------------------------------------------------------------------------------------
// Column-per-lane panel (-DQRK_WY_COLPANEL).  Lane (j = lane >> 2, q = lane & 3) holds column j of the panel, local rows 4 i + q
// (i < RQ = 8 MR) — its OWN column only, so the two things a Householder step must share travel through shared memory
// instead of a warp-wide reduction:
//   * the RAW pivot column (published by the four lanes that own it as soon as the previous step has updated it), its
//     partial tail norms, and the pivot row;
//   * nothing else: every lane forms the dot product of the raw tail with its own column over its own rows, and two
//     shuffles (xor 1, 2) finish it.  The scalar chain (rsqrt / reciprocal with Newton steps, all lanes redundantly) runs
//     beside those dot products: tau v^T a_j = -(dd a_Kj + t_j) / beta needs only the raw dot t_j.
// Lanes of the already finished columns j < K take part in the same dot product: their result is S_jK = v_j^T v_K, the
// entry of V^T V the T factor needs (dlarft), so T costs no extra pass.  One __syncwarp per step (the step buffers
// alternate).  Per step and lane: 2 RQ shared loads, 3 RQ FMAs, 2 shuffles, one scalar chain — against ONE batched
// 32-lane reduce-scatter (9 dependent shuffle rounds) + gather per step of the row-per-lane panel above.
// ---------------------------------------------------------------------------------------------------------------
template <int MR>
__device__ __noinline__ void wy_factor_panel_cpl(double* sA, int ld, int rp, int p, double* PB0, double* sT, double* scratch,
                                                 double* sTau, int lane) {
  constexpr int RQ = 8 * MR, TS = RQ + 2;
  constexpr int ROWB = 4 * (32 + 2);              // pivot row (8 columns) at ROWB, the 4 norm partials at ROWB + 8
  const int nrow = rp - p;                        // a multiple of 8
  const int nq = nrow >> 2;                       // valid rows per lane; the rest of a[] is zero and stays zero
  const int j = lane >> 2, q = lane & 3;
  double* sS = scratch + 2 * kWyStepBuf;          // 8 x 8, strictly upper part used
  WY_TRACE(0);
  double a[RQ];
  {
    const double* base = sA + (size_t)(p + j) * ld + p + q;
#pragma unroll
    for (int i = 0; i < RQ; i++) a[i] = (i < nq) ? base[4 * i] : 0.0;
  }
  __syncwarp();                                   // the scratch of an earlier use (previous panel, cooperative apply) is dead
  if (j == 0) {                                   // what step 0 reads: raw column 0, its tail norm partials ...
    double* dst = scratch + q * TS;
    double s0 = (q > 0) ? a[0] * a[0] : 0.0, s1 = a[1] * a[1], s2 = 0.0, s3 = 0.0;
    *reinterpret_cast<double2*>(dst) = make_double2(a[0], a[1]);
#pragma unroll
    for (int i = 2; i < RQ; i += 2) {
      *reinterpret_cast<double2*>(dst + i) = make_double2(a[i], a[i + 1]);
      if (i & 2) { s2 = fma(a[i], a[i], s2); s3 = fma(a[i + 1], a[i + 1], s3); }
      else { s0 = fma(a[i], a[i], s0); s1 = fma(a[i + 1], a[i + 1], s1); }
    }
    scratch[ROWB + 8 + q] = (s0 + s1) + (s2 + s3);
  }
  if (q == 0) scratch[ROWB + j] = a[0];           // ... and row 0 of the panel
  double* tauv = sTau + p;
  double myinv = 0.0;                             // 1 / (x0 - beta) of this lane's OWN column: the essential part is stored
                                                  // unscaled in a[] and scaled once, at write-back
  WY_TRACE(1);
#pragma unroll 1
  for (int K = 0; K < 8; K++) {
    const double* buf = scratch + (K & 1) * kWyStepBuf;
    const double* tailq = buf + q * TS;           // raw pivot column, rows 4 i + q
    __syncwarp();
    const double c0 = buf[ROWB + K], rowj = buf[ROWB + j];
    const double2 n01 = *reinterpret_cast<const double2*>(buf + ROWB + 8), n23 = *reinterpret_cast<const double2*>(buf + ROWB + 10);
    const double tailSq = (n01.x + n01.y) + (n23.x + n23.y);
    const bool m0 = q > K, m1 = 4 + q > K;        // rows q and 4 + q lie below the pivot row?
    // raw dot of the pivot column's tail with this lane's column (rows > K); rows >= 8 always take part
    double t0, t1, t2 = 0.0, t3 = 0.0;
    {
      const double2 tv = *reinterpret_cast<const double2*>(tailq);
      t0 = m0 ? tv.x * a[0] : 0.0;
      t1 = m1 ? tv.y * a[1] : 0.0;
    }
#pragma unroll
    for (int i = 2; i < RQ; i += 2) {
      const double2 tv = *reinterpret_cast<const double2*>(tailq + i);
      if (i & 2) { t2 = fma(tv.x, a[i], t2); t3 = fma(tv.y, a[i + 1], t3); }
      else { t0 = fma(tv.x, a[i], t0); t1 = fma(tv.y, a[i + 1], t1); }
    }
    double t = (t0 + t1) + (t2 + t3);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    const WyRefl h = wy_reflector(c0, tailSq);    // Eigen makeHouseholder; runs beside the dot products
    const double sj = -fma(h.dd, rowj, t) * h.ib; // tau v^T a_j
    const double add = (j > K) ? -(sj * h.inv) : 0.0;
    // columns j > K: a -= (tau v^T a_j) v, v = [1; inv * tail]; the others see add = 0 (a + 0 * tail: unchanged)
    {
      const double2 tv = *reinterpret_cast<const double2*>(tailq);
      a[0] = fma(tv.x, m0 ? add : 0.0, a[0]);
      a[1] = fma(tv.y, m1 ? add : 0.0, a[1]);
    }
#pragma unroll
    for (int i = 2; i < RQ; i += 2) {
      const double2 tv = *reinterpret_cast<const double2*>(tailq + i);
      a[i] = fma(tv.x, add, a[i]);
      a[i + 1] = fma(tv.y, add, a[i + 1]);
    }
    if (q == (K & 3)) {                           // the pivot row: R_Kj = a_Kj - tau v^T a_j (j > K), beta on the diagonal
      const double old = (K >> 2) ? a[1] : a[0];
      const double nv = (j > K) ? old - sj : (j == K) ? h.beta : old;
      if (K >> 2) a[1] = nv; else a[0] = nv;
    }
    if (j == K) myinv = h.inv;
    if (j < K && q == 0) sS[j * 8 + K] = myinv * fma(h.inv, t, rowj);   // S_jK = v_j[K] + v_j[K+1:]^T v_K[K+1:], v_j = myinv * a
    if (j == K && q == 0) tauv[K] = h.tau;
    if (K < 7) {                                  // publish what step K + 1 reads
      double* nb = scratch + ((K + 1) & 1) * kWyStepBuf;
      if (j == K + 1) {
        double* dst = nb + q * TS;
        double s0 = (q > K + 1) ? a[0] * a[0] : 0.0, s1 = (4 + q > K + 1) ? a[1] * a[1] : 0.0, s2 = 0.0, s3 = 0.0;
        *reinterpret_cast<double2*>(dst) = make_double2(a[0], a[1]);
#pragma unroll
        for (int i = 2; i < RQ; i += 2) {
          *reinterpret_cast<double2*>(dst + i) = make_double2(a[i], a[i + 1]);
          if (i & 2) { s2 = fma(a[i], a[i], s2); s3 = fma(a[i + 1], a[i + 1], s3); }
          else { s0 = fma(a[i], a[i], s0); s1 = fma(a[i + 1], a[i + 1], s1); }
        }
        nb[ROWB + 8 + q] = (s0 + s1) + (s2 + s3);
      }
      if (q == ((K + 1) & 3)) nb[ROWB + j] = ((K + 1) >> 2) ? a[1] : a[0];
    }
  }
  WY_TRACE(2);
  // packed columns back in place: R above / on the diagonal, the essential parts (scaled now) below; unit-lower diagonal tile of V
  {
    const double e0 = (q > j) ? a[0] * myinv : a[0], e1 = (4 + q > j) ? a[1] * myinv : a[1];
    double* base = sA + (size_t)(p + j) * ld + p + q;
    base[0] = e0; base[4] = e1;
#pragma unroll
    for (int i = 2; i < RQ; i++) if (i < nq) base[4 * i] = a[i] * myinv;
    PB0[j * 12 + q] = (q < j) ? 0.0 : (q == j) ? 1.0 : e0;
    PB0[j * 12 + 4 + q] = (4 + q < j) ? 0.0 : (4 + q == j) ? 1.0 : e1;
  }
  __syncwarp();
  // T (upper triangular, Q = I - V T V^T): T_kk = tau_k, T[0:k,k] = -tau_k T[0:k,0:k] S[0:k,k]   (LAPACK dlarft /
  // Eigen make_block_householder_triangular_factor); lane i owns row i
  if (lane < 8) {
    double Tr[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      double sum = 0.0;
#pragma unroll
      for (int m = 0; m < k; m++) sum = fma(Tr[m], sS[m * 8 + k], sum);
      const double tk = tauv[k];
      Tr[k] = (k < lane) ? 0.0 : (k == lane) ? tk : -tk * sum;
      sT[lane * 8 + k] = Tr[k];
    }
  }
  __syncwarp();
  WY_TRACE(5);
}


// ---------------------------------------------------------------------------------------------------------------
// Team panel (-DQRK_WY_TEAMPANEL, experiment): the column-per-lane panel spread over ALL WP warps of the CTA.  Thread
// (warp w, lane) = (column j = lane >> 2, sub-chunk s = 4 w + (lane & 3)) holds the rows NS i + s of column j (NS = 4 WP,
// 8 MR / WP rows per thread), so the panel's pivot rows 0..7 are all "row i = 0" of the sub-chunks s = 0..7.  Per step: the
// raw pivot column, its tail-norm partials (one per warp) and the pivot row are published in shared memory; every thread
// forms its dot-product partial, two shuffles reduce over the lane's four sub-chunks, the WP per-warp partials cross through
// shared memory; the scalar chain runs beside all that.  Two CTA barriers per step.  Every warp of the CTA must call it.
// ---------------------------------------------------------------------------------------------------------------
template <int MR, int WP>
__device__ __noinline__ void wy_factor_panel_team(double* sA, int ld, int rp, int p, double* PB0, double* sT, double* scratch,
                                                  double* sTau, int warp, int lane) {
  static_assert(WP == 2 || WP == 4, "pivot rows 0..7 must all be row 0 of a sub-chunk");
  constexpr int NS = 4 * WP, RQ = 8 * MR / WP, TS = RQ + 2;
  constexpr int ROWB = 16 * (16 + 2);             // pivot row (8) at ROWB, norm partials (4) at ROWB + 8, dot partials (4 x 8) at ROWB + 12
  static_assert(ROWB + 12 + 32 + 64 <= kWyScratch, "scratch too small for the team panel");
  const int nrow = rp - p;
  const int j = lane >> 2, q = lane & 3, sc = 4 * warp + q;
  double* sS = scratch + ROWB + 44;               // 8 x 8, strictly upper part used
  double* tail = scratch;                         // [NS][TS] raw pivot column
  double* rowb = scratch + ROWB;
  double* nrm = rowb + 8;
  double* dotp = rowb + 12;
  double a[RQ];
  {
    const double* base = sA + (size_t)(p + j) * ld + p + sc;
#pragma unroll
    for (int i = 0; i < RQ; i++) a[i] = (NS * i + sc < nrow) ? base[NS * i] : 0.0;
  }
  auto publish = [&](int Kn) {                    // what step Kn reads: raw column Kn, its tail-norm partial per warp, row Kn
    if (j == Kn) {
      double* dst = tail + sc * TS;
      double s0 = (sc > Kn) ? a[0] * a[0] : 0.0, s1 = 0.0;
      dst[0] = a[0];
#pragma unroll
      for (int i = 1; i < RQ; i++) {
        dst[i] = a[i];
        if (i & 1) s1 = fma(a[i], a[i], s1); else s0 = fma(a[i], a[i], s0);
      }
      double sn = s0 + s1;
      sn += __shfl_xor_sync(0xfu << (4 * Kn), sn, 1);
      sn += __shfl_xor_sync(0xfu << (4 * Kn), sn, 2);
      if (q == 0) nrm[warp] = sn;
    }
    if (sc == Kn) rowb[j] = a[0];
  };
  __syncthreads();                                // the scratch of an earlier use is dead
  publish(0);
  double* tauv = sTau + p;
  double myinv = 0.0;
  WY_TRACE(1);
#ifdef QRK_WY_TRACE
  long long wy_acc[6] = {0, 0, 0, 0, 0, 0};
  long long wy_last = clock64();
#endif
#pragma unroll 1
  for (int K = 0; K < 8; K++) {
    __syncthreads();
    WY_CLK(0);                                    // barrier A (publication of the previous step)
    const double c0 = rowb[K], rowj = rowb[j];
    double tailSq = nrm[0] + nrm[1];
    if (WP == 4) tailSq += nrm[2] + nrm[3];
    double tv[RQ];
    double t0, t1 = 0.0;
    {
      const double* tq = tail + sc * TS;
#pragma unroll
      for (int i = 0; i < RQ; i++) tv[i] = tq[i];
      t0 = (sc > K) ? tv[0] * a[0] : 0.0;
#pragma unroll
      for (int i = 1; i < RQ; i++) { if (i & 1) t1 = fma(tv[i], a[i], t1); else t0 = fma(tv[i], a[i], t0); }
    }
    double t = t0 + t1;
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    if (q == 0) dotp[warp * 8 + j] = t;
    WY_CLK(1);                                    // shared loads, dot partial, two shuffles
    const WyRefl h = wy_reflector(c0, tailSq);
#ifdef QRK_WY_TRACE
    if (h.tau == 123.456) wy_acc[5]++;            // keep the chain in front of the clock read
#endif
    WY_CLK(2);                                    // scalar chain
    __syncthreads();
    WY_CLK(3);                                    // barrier B (dot partials of all warps)
    t = dotp[j] + dotp[8 + j];
    if (WP == 4) t += dotp[16 + j] + dotp[24 + j];
    const double sj = -fma(h.dd, rowj, t) * h.ib;
    const double add = (j > K) ? -(sj * h.inv) : 0.0;
    {
      const double old = a[0];
      a[0] = fma(tv[0], (sc > K) ? add : 0.0, a[0]);
      if (sc == K) a[0] = (j > K) ? old - sj : (j == K) ? h.beta : old;
    }
#pragma unroll
    for (int i = 1; i < RQ; i++) a[i] = fma(tv[i], add, a[i]);
    if (j == K) myinv = h.inv;
    if (j < K && sc == 0) sS[j * 8 + K] = myinv * fma(h.inv, t, rowj);
    if (j == K && sc == 0) tauv[K] = h.tau;
    if (K < 7) publish(K + 1);
    WY_CLK(4);                                    // update + publication
  }
#ifdef QRK_WY_TRACE
  for (int i = 0; i < 5; i++) WY_TRACE_VAL(12 + i, wy_acc[i]);
#endif
  WY_TRACE(2);
  {
    const double e0 = (sc > j) ? a[0] * myinv : a[0];
    double* base = sA + (size_t)(p + j) * ld + p + sc;
    if (sc < nrow) base[0] = e0;
#pragma unroll
    for (int i = 1; i < RQ; i++) if (NS * i + sc < nrow) base[NS * i] = a[i] * myinv;
    if (sc < 8) PB0[j * 12 + sc] = (sc < j) ? 0.0 : (sc == j) ? 1.0 : e0;
  }
  __syncthreads();
  if (warp == 0 && lane < 8) {
    double Tr[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      double sum = 0.0;
#pragma unroll
      for (int m = 0; m < k; m++) sum = fma(Tr[m], sS[m * 8 + k], sum);
      const double tk = tauv[k];
      Tr[k] = (k < lane) ? 0.0 : (k == lane) ? tk : -tk * sum;
      sT[lane * 8 + k] = Tr[k];
    }
  }
}

// ---- apply panel p to one 8-column tile (or to the right-hand side), by ONE warp ---------------------------------
template <int MR>
__device__ __forceinline__ void wy_apply_panel(double* sA, double* sRhs, int ld, int rp, int p, int jt, bool is_rhs,
                                               const double* PB0, const double* sT, int lane) {
  constexpr int NT = 4 * MR;
  const int q = lane & 3, g = lane >> 2;
  const int k0 = wy_kappa(q, 0), k1 = wy_kappa(q, 1);
  const int rho = wy_kappa(g >> 1, g & 1);
  const int nt = (rp - p) >> 3;
  const bool valid = !is_rhs || g == 0;
  double* col = (is_rhs ? sRhs : sA + (size_t)(8 * jt + g) * ld) + p;
  double a0[NT], a1[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) {
    const bool ld_ok = valid && t < nt;
    a0[t] = ld_ok ? col[8 * t + k0] : 0.0;
    a1[t] = ld_ok ? col[8 * t + k1] : 0.0;
  }
  // W^T = A2^T V   (8 columns x 8 reflectors)
  const double* vcol = sA + (size_t)(p + g) * ld + p;
  double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0, e00 = 0.0, e01 = 0.0, e10 = 0.0, e11 = 0.0;   // four independent chains
#pragma unroll
  for (int t = 0; t < NT; t++) {
    if (t < nt) {
      const double v0 = (t == 0) ? PB0[g * 12 + k0] : vcol[8 * t + k0];
      const double v1 = (t == 0) ? PB0[g * 12 + k1] : vcol[8 * t + k1];
      if (t & 1) { dmma884(e00, e01, a0[t], v0); dmma884(e10, e11, a1[t], v1); }
      else { dmma884(c00, c01, a0[t], v0); dmma884(c10, c11, a1[t], v1); }
    }
  }
  const double w0 = (c00 + c10) + (e00 + e10), w1 = (c01 + c11) + (e01 + e11);
  // W2^T = W^T T ; output slots 2q+e hold reflector q + 4e
  const int pg = (g >> 1) + 4 * (g & 1);
  double d0 = 0.0, d1 = 0.0;
  dmma884(d0, d1, w0, sT[(2 * q) * 8 + pg]);
  dmma884(d0, d1, w1, sT[(2 * q + 1) * 8 + pg]);
  d0 = -d0; d1 = -d1;
  // A2^T -= W2^T V^T
  const double* vq0 = sA + (size_t)(p + q) * ld + p + rho;
  const double* vq1 = sA + (size_t)(p + q + 4) * ld + p + rho;
#pragma unroll
  for (int t = 0; t < NT; t++) {
    if (t < nt) {
      const double b0 = (t == 0) ? PB0[q * 12 + rho] : vq0[8 * t];
      const double b1 = (t == 0) ? PB0[(q + 4) * 12 + rho] : vq1[8 * t];
      dmma884(a0[t], a1[t], d0, b0);
      dmma884(a0[t], a1[t], d1, b1);
      if (valid) { col[8 * t + k0] = a0[t]; col[8 * t + k1] = a1[t]; }
    }
  }
}


// ---- apply panel p to the NEXT panel's tile, by ALL W warps of the CTA (the look-ahead tile is on the critical path) ----
// Warp w takes the row tiles t = w, w + W, ...: partial W^T = A2^T V over its rows -> shared memory -> every warp sums the W
// partials, multiplies by T and updates its own row tiles.  One __syncthreads inside; every warp of the CTA must call it.
template <int MR, int W>
__device__ __forceinline__ void wy_apply_panel_coop(double* sA, int ld, int rp, int p, int jt, const double* PB0, const double* sT,
                                                    double* sWpart, int warp, int lane) {
  constexpr int NTW = (4 * MR + W - 1) / W;
  const int q = lane & 3, g = lane >> 2;
  const int k0 = wy_kappa(q, 0), k1 = wy_kappa(q, 1);
  const int rho = wy_kappa(g >> 1, g & 1);
  const int nt = (rp - p) >> 3;
  double* col = sA + (size_t)(8 * jt + g) * ld + p;
  double a0[NTW], a1[NTW];
#pragma unroll
  for (int u = 0; u < NTW; u++) {
    const int t = warp + u * W;
    a0[u] = (t < nt) ? col[8 * t + k0] : 0.0;
    a1[u] = (t < nt) ? col[8 * t + k1] : 0.0;
  }
  const double* vcol = sA + (size_t)(p + g) * ld + p;
  double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll
  for (int u = 0; u < NTW; u++) {
    const int t = warp + u * W;
    if (t < nt) {
      const double v0 = (t == 0) ? PB0[g * 12 + k0] : vcol[8 * t + k0];
      const double v1 = (t == 0) ? PB0[g * 12 + k1] : vcol[8 * t + k1];
      dmma884(c00, c01, a0[u], v0);
      dmma884(c10, c11, a1[u], v1);
    }
  }
  *reinterpret_cast<double2*>(sWpart + warp * 64 + 2 * lane) = make_double2(c00 + c10, c01 + c11);
  __syncthreads();
  double w0 = 0.0, w1 = 0.0;
#pragma unroll
  for (int w = 0; w < W; w++) {
    const double2 pw = *reinterpret_cast<const double2*>(sWpart + w * 64 + 2 * lane);
    w0 += pw.x; w1 += pw.y;
  }
  const int pg = (g >> 1) + 4 * (g & 1);
  double d0 = 0.0, d1 = 0.0;
  dmma884(d0, d1, w0, sT[(2 * q) * 8 + pg]);
  dmma884(d0, d1, w1, sT[(2 * q + 1) * 8 + pg]);
  d0 = -d0; d1 = -d1;
  const double* vq0 = sA + (size_t)(p + q) * ld + p + rho;
  const double* vq1 = sA + (size_t)(p + q + 4) * ld + p + rho;
#pragma unroll
  for (int u = 0; u < NTW; u++) {
    const int t = warp + u * W;
    if (t < nt) {
      const double b0 = (t == 0) ? PB0[q * 12 + rho] : vq0[8 * t];
      const double b1 = (t == 0) ? PB0[(q + 4) * 12 + rho] : vq1[8 * t];
      dmma884(a0[u], a1[u], d0, b0);
      dmma884(a0[u], a1[u], d1, b1);
      col[8 * t + k0] = a0[u]; col[8 * t + k1] = a1[u];
    }
  }
}

// ---- kernel: grid.x = blocks of this size class -----------------------------------------------------------------
// Measured on B200 (profiles/r01_wy_gram_panel_experiment.txt): the Gram-downdated panel shortens a column step from
// ~1100 to ~830 cycles but its two wide reductions (G: 36 sums, S: 28 sums) cost 2.2-3.1 k cycles each, so config 5 runs
// at 7.3 ms instead of 6.6 ms: the per-column all-reduce stays the default; -DQRK_WY_GRAM selects the other variant.
// Measured on B200, config 5 per class (profiles/r02_wy_panel_variants.md): with the cooperative look-ahead apply and the tile
// queue below, the row-per-lane panel gives 6.25 ms, the column-per-lane panel 6.75 ms (its step has no warp-wide reduction
// but ~315 instructions, two shared-memory passes over the pivot column and 4 RQ FP64 instructions: 1.65 k cycles against
// 1.1 k), the round-1 schedule 6.63 ms.  The row-per-lane panel stays the default.
#if defined(QRK_WY_GRAM)
#define WY_FACTOR_PANEL wy_factor_panel_gram   // row-per-lane, one Gram reduction per panel (experiment)
#elif defined(QRK_WY_COLPANEL)
#define WY_FACTOR_PANEL wy_factor_panel_cpl    // column-per-lane, no warp-wide reduction (experiment)
#else
#define WY_FACTOR_PANEL wy_factor_panel        // row-per-lane, one all-reduce per column (default)
#endif

// resident CTAs per SM the register allocation must leave room for (shared memory allows about as many)
#ifndef QRK_WY_WARPS_MR3
#define QRK_WY_WARPS_MR3 12
#endif
#ifndef QRK_WY_WARPS_MR2
#define QRK_WY_WARPS_MR2 16
#endif
#ifndef QRK_WY_WARPS_MR1
#define QRK_WY_WARPS_MR1 20
#endif
__host__ __device__ constexpr int wy_min_ctas(int mr, int w) { return (mr >= 3 ? QRK_WY_WARPS_MR3 : mr == 2 ? QRK_WY_WARPS_MR2 : QRK_WY_WARPS_MR1) / w; }

template <int MR, int W, bool SOLVE>
__global__ void __launch_bounds__(32 * W, wy_min_ctas(MR, W))
bd_wy_factor_kernel(BlockIndex bi, const int* __restrict__ ids, const double* A_in, double* packed,
                    double* __restrict__ tau_out, const double* __restrict__ b, double* __restrict__ x) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const long long blk = ids ? ids[blockIdx.x] : blockIdx.x;
  int r, c;
  long long vo, ro, co;
  bi.get(blk, r, c, vo, ro, co);
  const WyGeom G(r, c);
  const int rp = G.rp, cp = G.cp, ld = G.ld;
  constexpr int T = 32 * W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  double* sA = reinterpret_cast<double*>(smem_raw);
  double* sRhs = sA + (size_t)cp * ld;
  double* sPB = sRhs + ld;
  double* sT = sPB + 2 * kWyPB;
  double* sS = sT + 2 * 64;
  double* sTau = sS + kWyScratch;
  double* sRd = sTau + cp;

  // panel pi is factored by warp (pi + rot) % W.  The CTAs resident on one SM run in lock step and the panel warp is
  // serial and FP64-heavy: if they all used warp 0 the three or more panel chains would share ONE sub-partition's
  // FP64 pipe.  A per-SM ticket gives consecutive CTAs of an SM consecutive rotations.
  __shared__ int s_rot;
  if (W > 1) {
    if (tid == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      s_rot = (int)(atomicAdd(&g_wy_sm_ticket[smid & 255], 1u) % W);
    }
  }
  // ---- stage the block: contiguous column-major r x c in HBM -> padded columns in shared memory ----
  const double* gA = A_in + vo;
  const bool vec2 = ((r & 1) == 0) && ((reinterpret_cast<uintptr_t>(gA) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(packed + vo) & 15) == 0);
  if (vec2) {
    const int rc = r >> 1;                         // 16-byte chunks per column
    int j = tid / rc, i = tid - j * rc;
    while (j < c) {
      cp_async16(sA + (size_t)j * ld + 2 * i, gA + (size_t)j * r + 2 * i);
      i += T;
      while (i >= rc) { i -= rc; j++; }
    }
    cp_async_commit();
  } else {
    int j = tid / r, i = tid - j * r;
    while (j < c) {
      sA[(size_t)j * ld + i] = gA[(size_t)j * r + i];
      i += T;
      while (i >= r) { i -= r; j++; }
    }
  }
  {
    const int padr = rp - r;                      // zero rows below the block: they stay zero under every reflector
    for (int e = tid; e < padr * c; e += T) { const int j = e / padr, i = e - j * padr; sA[(size_t)j * ld + r + i] = 0.0; }
    for (int e = tid; e < (cp - c) * rp; e += T) { const int j = e / rp, i = e - j * rp; sA[(size_t)(c + j) * ld + i] = 0.0; }
    if (SOLVE) for (int i = tid; i < rp; i += T) sRhs[i] = (i < r) ? b[ro + i] : 0.0;
  }
  if (vec2) cp_async_wait<0>();
  __syncthreads();

  WY_TRACE(11);
  const int P = cp >> 3;
  const int n_tiles = P + (SOLVE ? 1 : 0);
  const int rot = (W > 1) ? s_rot : 0;
  __shared__ int s_next_tile;
#ifdef QRK_WY_TEAMPANEL
  if constexpr (W == 2 || W == 4) {
    // experiment: every panel by the whole CTA (wy_factor_panel_team), the tiles by the whole CTA from the queue: no overlap
    wy_factor_panel_team<MR, W>(sA, ld, rp, 0, sPB, sT, sS, sTau, warp, lane);
    __syncthreads();
    for (int pi = 0; pi < P; pi++) {
      const int p = 8 * pi, buf = pi & 1;
      const bool has_next = pi + 1 < P;
      int first = pi + 1;
      if (has_next) {
        WY_TRACE(6);
        wy_apply_panel_coop<MR, W>(sA, ld, rp, p, pi + 1, sPB + buf * kWyPB, sT + buf * 64, sS, warp, lane);
        WY_TRACE(7);
        first = pi + 2;
      }
      if (tid == 0) s_next_tile = first;
      __syncthreads();
      WY_TRACE(3);
      for (;;) {
        int jt = 0;
        if (lane == 0) jt = atomicAdd(&s_next_tile, 1);
        jt = __shfl_sync(0xffffffffu, jt, 0);
        if (jt >= n_tiles) break;
        wy_apply_panel<MR>(sA, sRhs, ld, rp, p, jt, jt == P, sPB + buf * kWyPB, sT + buf * 64, lane);
        WY_TRACE(4);
      }
      WY_TRACE(8);
      __syncthreads();
      WY_TRACE(9);
      if (has_next) {
        wy_factor_panel_team<MR, W>(sA, ld, rp, p + 8, sPB + (buf ^ 1) * kWyPB, sT + (buf ^ 1) * 64, sS, sTau, warp, lane);
        __syncthreads();
      }
    }
  } else
#endif
  {
#ifdef QRK_WY_PINPANEL
  if (warp == 0) WY_FACTOR_PANEL<MR>(sA, ld, rp, 0, sPB, sT, sS, sTau, lane);
#else
  if (warp == rot) WY_FACTOR_PANEL<MR>(sA, ld, rp, 0, sPB, sT, sS, sTau, lane);
#endif
  __syncthreads();
  for (int pi = 0; pi < P; pi++) {
    const int p = 8 * pi, buf = pi & 1;
    const bool has_next = pi + 1 < P;
    if (W == 1) {
      for (int jt = pi + 1; jt < n_tiles; jt++) {
        wy_apply_panel<MR>(sA, sRhs, ld, rp, p, jt, jt == P, sPB + buf * kWyPB, sT + buf * 64, lane);
        if (has_next && jt == pi + 1) {
          __syncwarp();
          WY_FACTOR_PANEL<MR>(sA, ld, rp, p + 8, sPB + (buf ^ 1) * kWyPB, sT + (buf ^ 1) * 64, sS, sTau, lane);
        }
      }
      __syncwarp();
      continue;
    }
    // (a) the look-ahead tile (the next panel's columns) by all warps, rows split; (b) its owner factors it at once while the
    // other warps take the remaining tiles (and the right-hand side) from a queue, the owner joins them when it is done
    int first = pi + 1;
    if (has_next) {
      WY_TRACE(6);
      wy_apply_panel_coop<MR, W>(sA, ld, rp, p, pi + 1, sPB + buf * kWyPB, sT + buf * 64, sS, warp, lane);
      WY_TRACE(7);
      first = pi + 2;
    }
    if (tid == 0) s_next_tile = first;
    __syncthreads();
#ifdef QRK_WY_PINPANEL
    // experiment: every panel of every resident CTA on warp 0 (one SM sub-partition runs the latency-bound scalar chains, the
    // DMMA tile streams stay on the other three); QRK_WY_PINPANEL=2: warp 0 takes no tiles either
    const int onext = 0;
#else
    const int onext = (pi + 1 + rot) % W;
#endif
    if (has_next && warp == onext)
      WY_FACTOR_PANEL<MR>(sA, ld, rp, p + 8, sPB + (buf ^ 1) * kWyPB, sT + (buf ^ 1) * 64, sS, sTau, lane);
    WY_TRACE(3);
#if defined(QRK_WY_PINPANEL) && QRK_WY_PINPANEL == 2
    if (warp != 0 || !has_next)
#endif
    for (;;) {
      int jt = 0;
      if (lane == 0) jt = atomicAdd(&s_next_tile, 1);
      jt = __shfl_sync(0xffffffffu, jt, 0);
      if (jt >= n_tiles) break;
      wy_apply_panel<MR>(sA, sRhs, ld, rp, p, jt, jt == P, sPB + buf * kWyPB, sT + buf * 64, lane);
      WY_TRACE(4);
    }
    WY_TRACE(8);
    __syncthreads();
    WY_TRACE(9);
  }
  }

  // ---- epilogue: x = R^-1 (Q^T b)[0:c] (warp 0, y in registers), tau, packed factors ----
  if (SOLVE) {
    for (int j = tid; j < c; j += T) sRd[j] = 1.0 / sA[(size_t)j * ld + j];
    __syncthreads();
    if (warp == 0) {
      double y[MR];
#pragma unroll
      for (int m = 0; m < MR; m++) y[m] = (lane + 32 * m < c) ? sRhs[lane + 32 * m] : 0.0;
      for (int j = c - 1; j >= 0; --j) {
        const int mj = j >> 5;
        double ym = y[0];
#pragma unroll
        for (int m = 1; m < MR; m++) if (mj == m) ym = y[m];
        const double yj = __shfl_sync(0xffffffffu, ym, j & 31) * sRd[j];
        const double* cj = sA + (size_t)j * ld;
#pragma unroll
        for (int m = 0; m < MR; m++) {
          const int i = lane + 32 * m;
          if (i < j) y[m] = fma(-cj[i], yj, y[m]);
          else if (i == j) y[m] = yj;
        }
      }
#pragma unroll
      for (int m = 0; m < MR; m++) if (lane + 32 * m < c) x[co + lane + 32 * m] = y[m];
    }
  }
  WY_TRACE(10);
  for (int j = tid; j < c; j += T) tau_out[co + j] = sTau[j];
  double* gP = packed + vo;
  if (vec2) {
    const int rc = r >> 1;
    int j = tid / rc, i = tid - j * rc;
    while (j < c) {
      *reinterpret_cast<double2*>(gP + (size_t)j * r + 2 * i) = *reinterpret_cast<const double2*>(sA + (size_t)j * ld + 2 * i);
      i += T;
      while (i >= rc) { i -= rc; j++; }
    }
  } else {
    int j = tid / r, i = tid - j * r;
    while (j < c) {
      gP[(size_t)j * r + i] = sA[(size_t)j * ld + i];
      i += T;
      while (i >= r) { i -= r; j++; }
    }
  }
}

}  // namespace qrk
