// banded.cuh — block-banded QR (BandedBlockedSparseQR, reference src/QRKit/BandedBlockedSparseQR.h:443-519):
// a strictly sequential sliding-window Householder recurrence (window recurrence :494-507), hence ONE GPU,
// one SM, and a latency-bound kernel by construction (SURVEY §7 hard part 5).
//
// Input: nb block rows; block row k is a dense BR x BC slab at rows [k*BR, (k+1)*BR), columns
// [k*S, k*S + BC), S = BC - OV (fromBlockBandedPattern(rows, cols, BR, BC, OV), SparseQRUtils.h:274-302),
// stored as block-COO slabs (column-major BR x BC, back to back).
//
// Window k = [carry (OV rows, upper triangular, left over from window k-1) ; slab k] is (OV+BR) x BC.
// A full Householder QR of the window makes its first S rows final rows of R (rows k*S .. k*S+S-1, BC
// entries each), annihilates OV+BR-BC rows, and leaves rows S..BC-1 as the next carry — the same
// recurrence as the reference (which first merges block rows into larger windows; R is unique up to row
// signs for a fixed column order, so the blocking is free).
//
// One warp, the whole window in registers, COLUMNS ACROSS LANES: lane j owns column j of the window (lane
// BC owns the right-hand side), so every v^T a_j dot product and every rank-1 update is lane-local with no
// reduction; the only cross-lane traffic is the broadcast of each reflector (<= BR values) from its lane.
// Because the carry is upper triangular, reflector c < OV touches only the pivot entry and the BR slab
// rows.  The next slab is prefetched into registers while the current window is factored.
// Outputs: band R (n_cols x BC, row g holds columns [w(g)*S, w(g)*S+BC)), y = (Q^T b) thin part, the
// packed windows (reflectors below the diagonal, in place over the slabs) and tau for later Q^T applications.
#pragma once
#include "common.cuh"

namespace qrk {

template <int BR, int BC, int OV>
struct BandedCfg {
  static constexpr int S = BC - OV;          // column step = rows finalised per window
  static constexpr int M = OV + BR;          // window rows
  static constexpr int DIE = M - BC;         // rows annihilated per window
  static_assert(BC + 1 <= 32, "one lane per window column plus the right-hand side");
  static_assert(OV >= 0 && OV < BC && M >= BC, "window must have at least as many rows as columns");
};

// window row r (0 <= r < M) of this lane's column lives in cw[r] (r < OV) or bw[r - OV]
#define QRK_WROW(r) ((r) < OV ? cw[(r) < OV ? (r) : 0] : bw[(r) >= OV ? (r) - OV : 0])

template <int BR, int BC, int OV>
__global__ void __launch_bounds__(32, 1)
banded_factor_kernel(const double* A_in, double* packed, double* __restrict__ tau_out, double* __restrict__ rband,
                     const double* __restrict__ b, double* __restrict__ y, double* __restrict__ ycomp, long long nb, int last_cols) {
  using G = BandedCfg<BR, BC, OV>;
  constexpr int S = G::S, M = G::M;
  const int lane = threadIdx.x;
  const bool is_col = lane < BC, is_rhs = (lane == BC) && (b != nullptr);
  __shared__ __align__(16) double sv[2 * BR];
  double cw[OV > 0 ? OV : 1], bw[BR], nxt[BR], tau_mine = 0.0;
#pragma unroll
  for (int i = 0; i < (OV > 0 ? OV : 1); i++) cw[i] = 0.0;

  auto load_slab = [&](long long k, double (&dst)[BR]) {
    if (is_col) {
      const double* src = A_in + (k * BC + lane) * (long long)BR;
#pragma unroll
      for (int i = 0; i < BR; i++) dst[i] = src[i];
    } else if (is_rhs) {
#pragma unroll
      for (int i = 0; i < BR; i++) dst[i] = b[k * BR + i];
    } else {
#pragma unroll
      for (int i = 0; i < BR; i++) dst[i] = 0.0;
    }
  };
  load_slab(0, nxt);

  for (long long k = 0; k < nb; k++) {
#pragma unroll
    for (int i = 0; i < BR; i++) bw[i] = nxt[i];
    if (k + 1 < nb) load_slab(k + 1, nxt);       // in flight while this window is factored
    const int ncols_w = (k == nb - 1) ? last_cols : BC;   // the last slab may be narrower (fromBlockBandedPattern, SparseQRUtils.h:284)

    // One column step = ONE dependent chain: [own norm^2 | raw dot with the published tail] -> scalars -> 2 broadcasts ->
    // update.  Lane c publishes its RAW tail t (no scaling), so every other lane's dot product t^T a_j runs concurrently
    // with lane c's rsqrt / reciprocal chain instead of behind it:  v = [1; inv t],  tau v^T a_j = tau (a_cj + inv t^T a_j)
    // = w_j,  a_j -= (w_j inv) t.  Lane c keeps its raw tail and its inv; the essential parts are scaled once per window.
    double inv_mine = 1.0;
#pragma unroll
    for (int c = 0; c < BC; c++) {
      if (c >= ncols_w) continue;                  // columns beyond the matrix: no reflector
      const int P0 = (c < OV) ? 0 : c - OV + 1;   // first participating slab row (compile time after unrolling)
      double* vb = sv + (c & 1) * BR;              // the buffer alternates with the column parity, one __syncwarp per column
      if (lane == c) {
#pragma unroll
        for (int i = 0; i < BR; i++) if (i >= P0) vb[i] = bw[i];
      }
      // ---- reflector scalars of the lane's own column (every lane runs the arithmetic, lane c's result is used)
      double tq[4] = {0.0, 0.0, 0.0, 0.0};          // four partial sums: the dependent FMA chain is BR/4 long, not BR
#pragma unroll
      for (int i = 0; i < BR; i++) if (i >= P0) tq[i & 3] = fma(bw[i], bw[i], tq[i & 3]);
      const double tailSq = (tq[0] + tq[1]) + (tq[2] + tq[3]);
      double pv = QRK_WROW(c);
      double beta, inv, tau;
      householder_scalars(pv, tailSq, P0 >= BR, beta, inv, tau);
      __syncwarp();
      // ---- raw dot product with the published tail (independent of the scalar chain above)
      double t[BR];
      double dq[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int i = 0; i < BR; i++) {
        if (i >= P0) {
          t[i] = vb[i];
          dq[i & 3] = fma(t[i], bw[i], dq[i & 3]);
        }
      }
      const double dot = (dq[0] + dq[1]) + (dq[2] + dq[3]);
      const double tau_c = __shfl_sync(0xffffffffu, tau, c);
      const double inv_c = __shfl_sync(0xffffffffu, inv, c);
      const bool upd = lane > c;                  // columns right of c and the right-hand side
      const double w = upd ? tau_c * fma(inv_c, dot, pv) : 0.0;
      const double z = w * inv_c;
      pv -= w;
      if (lane == c) { pv = beta; tau_mine = tau; inv_mine = inv; }
      if (c < OV) cw[c < OV ? c : 0] = pv; else bw[c >= OV ? c - OV : 0] = pv;
#pragma unroll
      for (int i = 0; i < BR; i++) if (i >= P0) bw[i] = fma(-z, t[i], bw[i]);   // lane c: z = 0, keeps its raw tail
    }
    // essential parts (LAPACK packing): lane's own column below its pivot, scaled by its 1/(x0 - beta)
    {
      const int p0_lane = (lane < OV) ? 0 : lane - OV + 1;
#pragma unroll
      for (int i = 0; i < BR; i++) if (i >= p0_lane) bw[i] *= inv_mine;
    }

    // ---- outputs of this window
    const bool last = (k == nb - 1);
    if (is_col) {
      double* dst = packed + (k * BC + lane) * (long long)BR;
#pragma unroll
      for (int i = 0; i < BR; i++) dst[i] = bw[i];
      tau_out[k * BC + lane] = tau_mine;
    }
#pragma unroll
    for (int r = 0; r < BC; r++) {
      if (last ? (r < last_cols) : (r < S)) {
        const double val = QRK_WROW(r);
        const long long g = k * S + r;
        if (is_col) rband[g * BC + lane] = (lane >= r) ? val : 0.0;
        else if (is_rhs) y[g] = val;
      }
    }
    if (is_rhs && ycomp) {
#pragma unroll
      for (int r = BC; r < M; r++) ycomp[k * (M - BC) + (r - BC)] = QRK_WROW(r);
    }
    // ---- next carry: window rows S..BC-1, columns shifted left by S (the right-hand side lane keeps its own)
    if (OV > 0) {
      double ncw[OV > 0 ? OV : 1];
#pragma unroll
      for (int i = 0; i < OV; i++) {
        const double val = QRK_WROW(S + i);
        const double shifted = __shfl_down_sync(0xffffffffu, val, S);
        ncw[i] = (lane == BC) ? val : ((lane < BC - S && lane >= i) ? shifted : 0.0);   // keep the carry upper triangular
      }
#pragma unroll
      for (int i = 0; i < OV; i++) cw[i] = ncw[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Q^T applied to a new right-hand side on a stored factorisation (packed windows + tau): the same window
// sweep with only the vector; lane i owns slab row i, the OV carried entries are replicated in every lane.
// ---------------------------------------------------------------------------------------------
template <int BR, int BC, int OV>
__global__ void __launch_bounds__(32, 1)
banded_apply_qt_kernel(const double* __restrict__ packed, const double* __restrict__ tau_in, const double* __restrict__ b,
                       double* __restrict__ y, double* __restrict__ ycomp, long long nb, int last_cols) {
  using G = BandedCfg<BR, BC, OV>;
  constexpr int S = G::S, M = G::M;
  static_assert(BR <= 32, "one lane per slab row");
  const int lane = threadIdx.x;
  double cy[OV > 0 ? OV : 1];
#pragma unroll
  for (int i = 0; i < (OV > 0 ? OV : 1); i++) cy[i] = 0.0;
  for (long long k = 0; k < nb; k++) {
    double bi = (lane < BR) ? b[k * BR + lane] : 0.0;
    double vv[BC], tt[BC];
#pragma unroll
    for (int c = 0; c < BC; c++) {               // independent of the chain: issued up front
      vv[c] = (lane < BR) ? packed[(k * BC + c) * (long long)BR + lane] : 0.0;
      tt[c] = tau_in[k * BC + c];
    }
    const int ncols_w = (k == nb - 1) ? last_cols : BC;
#pragma unroll
    for (int c = 0; c < BC; c++) {
      if (c >= ncols_w) continue;
      const int P0 = (c < OV) ? 0 : c - OV + 1;
      double part = (lane >= P0 && lane < BR) ? vv[c] * bi : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      const double pivot = (c < OV) ? cy[c < OV ? c : 0] : __shfl_sync(0xffffffffu, bi, c >= OV ? c - OV : 0);
      const double w = tt[c] * (pivot + part);
      if (c < OV) cy[c < OV ? c : 0] -= w;
      else if (lane == c - OV) bi -= w;
      if (lane >= P0 && lane < BR) bi = fma(-vv[c], w, bi);
    }
    const bool last = (k == nb - 1);
#pragma unroll
    for (int r = 0; r < BC; r++) {
      if (last ? (r < last_cols) : (r < S)) {
        const double val = (r < OV) ? cy[r < OV ? r : 0] : __shfl_sync(0xffffffffu, bi, r >= OV ? r - OV : 0);
        if (lane == 0) y[k * S + r] = val;
      }
    }
    if (ycomp) {
#pragma unroll
      for (int r = BC; r < M; r++) {
        const double val = __shfl_sync(0xffffffffu, bi, r - OV);
        if (lane == 0) ycomp[k * (M - BC) + (r - BC)] = val;
      }
    }
    if (OV > 0) {
      double ncy[OV > 0 ? OV : 1];
#pragma unroll
      for (int i = 0; i < OV; i++)
        ncy[i] = (S + i < OV) ? cy[(S + i < OV) ? S + i : 0] : __shfl_sync(0xffffffffu, bi, (S + i >= OV) ? S + i - OV : 0);
#pragma unroll
      for (int i = 0; i < OV; i++) cy[i] = ncy[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Back substitution with the band R, from the bottom (R.topLeftCorner(rank, rank).triangularView<Upper>().solve,
// BandedBlockedSparseQR.h:299-304).  One warp, 32 rows per chunk, lane l owns row g = 32 q + l.  Column-oriented: as soon as
// x_t is known every lane subtracts its R(g, t) x_t, so the dependent chain per row is ONE FMA and ONE broadcast shuffle
// (the rows are pre-scaled by 1/R(g,g)); the coefficients sit in registers indexed by the COLUMN (compile time after
// unrolling), fetched from a cp.async ring of band rows that runs three chunks ahead of the chain.
// Band row g holds columns [w(g) S, w(g) S + BC), w(g) = min(g / S, nb - 1)  (banded_factor_kernel's rband layout).
// ---------------------------------------------------------------------------------------------
template <int BC, int OV>
__global__ void __launch_bounds__(32, 1)
banded_backsolve_kernel(const double* __restrict__ rband, const double* __restrict__ y, double* __restrict__ x, long long nb,
                        int last_cols) {
  constexpr int S = BC - OV, CH = 32, NST = 4, NA = BC - 1;
  static_assert(BC % 2 == 0, "16-byte cp.async of whole band rows");
  __shared__ __align__(16) double ring[NST][CH * BC];
  __shared__ double yring[NST][CH];
  const int lane = threadIdx.x;
  const long long n_cols = (nb - 1) * S + last_cols;
  const long long nchunks = (n_cols + CH - 1) / CH;

  auto prefetch = [&](long long q) {
    if (q >= 0) {
      const int st = (int)(q % NST);
      const long long r0 = q * CH;
      const int valid = (int)((n_cols - r0 < CH) ? (n_cols - r0) : CH);
      const double* src = rband + r0 * BC;
      for (int i = lane; i < valid * BC / 2; i += 32) cp_async16(&ring[st][2 * i], src + 2 * i);
      if (lane < valid) cp_async8(&yring[st][lane], y + r0 + lane);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int i = 0; i < NST - 1; i++) prefetch(nchunks - 1 - i);

  double xprev = 0.0;                             // x of the chunk below (columns 32 (q+1) + lane)
  for (long long q = nchunks - 1; q >= 0; --q) {
    prefetch(q - (NST - 1));
    cp_async_wait<NST - 1>();
    __syncwarp();
    const int st = (int)(q % NST);
    const long long g = q * CH + lane;
    const bool valid = g < n_cols;
    const long long w = (g / S < nb - 1) ? g / S : nb - 1;
    const int shift = (int)(q * CH - w * S);      // band offset of column 32 q + t is t + shift
    const int ncw = (w == nb - 1) ? last_cols : BC;
    const double* row = &ring[st][lane * BC];
    const double diag = valid ? row[lane + shift] : 1.0;
    const double dinv = valid ? 1.0 / diag : 0.0;
    double c[CH], a[NA > 0 ? NA : 1];
#pragma unroll
    for (int t = 0; t < CH; t++) {
      const int off = t + shift;
      const bool use = valid && (t > lane) && (off < ncw);
      c[t] = use ? row[use ? off : 0] * dinv : 0.0;
    }
#pragma unroll
    for (int u = 0; u < NA; u++) {
      const int off = CH + u + shift;
      const bool use = valid && (off < ncw);
      a[u] = use ? row[use ? off : 0] * dinv : 0.0;
    }
    double s = valid ? yring[st][lane] * dinv : 0.0;
#pragma unroll
    for (int u = 0; u < NA; u++) s = fma(-a[u], __shfl_sync(0xffffffffu, xprev, u), s);
#pragma unroll
    for (int t = CH - 1; t >= 0; --t) {
      const double xt = __shfl_sync(0xffffffffu, s, t);    // lane t's row is complete: s = x_t
      s = fma(-c[t], xt, s);                                // c[t] = 0 for lanes >= t
    }
    if (valid) x[g] = s;
    xprev = s;
    __syncwarp();                                  // the stage is refilled by the next iteration's prefetch
  }
}

#undef QRK_WROW

}  // namespace qrk
