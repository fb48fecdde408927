// banded.cuh — block-banded QR (BandedBlockedSparseQR, reference src/QRKit/BandedBlockedSparseQR.h:443-519).
//
// Input: nb block rows; block row k is a dense BR x BC slab at rows [k*BR, (k+1)*BR), columns
// [k*S, k*S + BC), S = BC - OV (fromBlockBandedPattern(rows, cols, BR, BC, OV), SparseQRUtils.h:274-302),
// stored as block-COO slabs (column-major BR x BC, back to back).
//
// The reference's window recurrence (:494-507) is a chain of nb dependent window QRs: window k = [carry (OV rows, upper
// triangular, left over from window k-1) ; slab k] is (OV+BR) x BC, its QR makes S rows of R final and leaves the next
// carry.  R is unique up to row signs for a fixed column order, so the rows may be eliminated in any order.
// Two phases:
//   1. banded_factor_kernel, PARALLEL over groups of G consecutive slabs: the same window recurrence, started from a zero
//      carry in every group, one warp per group.  Group g ends with its own band triangle R_g of W = (G-1) S + BC rows
//      (the last OV of them overlap the next group's first OV columns).  All the O(rows) elimination work is here.
//   2. banded_chase_kernel, ONE warp, sequential: the OV-row triangle handed over by group g-1 is merged into R_g by chasing
//      it down the band — one Householder reflector per column with a tail of only OV entries — which finalises S rows of R
//      per window with S column steps instead of BC, and hands the last OV rows on to group g+1.  This is the only
//      sequential part: n_cols + (groups-1) OV short steps instead of nb BC long ones.
// In both kernels a lane owns one COLUMN of the window (lane BC owns the right-hand side), so every dot product and every
// rank-1 update is lane-local; the only cross-lane traffic per column step is the broadcast of the reflector through shared
// memory and two scalars by shuffle.
// Outputs: band R (n_cols x BC, row g holds columns [w(g)*S, w(g)*S+BC)), y = (Q^T b) thin part; for later Q^T / Q
// applications the packed windows + tau of phase 1 and the chase reflectors (OV essentials + tau per step) of phase 2.
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
banded_factor_kernel(const double* A_in, double* packed, double* __restrict__ tau_out, double* __restrict__ gband,
                     const double* __restrict__ b, double* __restrict__ gy, long long nb_total, int last_cols, int group) {
  using G = BandedCfg<BR, BC, OV>;
  constexpr int S = G::S;
  const int lane = threadIdx.x;
  // this warp's group of slabs [k0, k0 + nb) and its slice of the group-band buffers (W rows per group)
  const long long k0 = (long long)blockIdx.x * group;
  const long long nb = (nb_total - k0 < group) ? (nb_total - k0) : group;
  const long long W = (long long)(group - 1) * S + BC;
  A_in += k0 * BC * (long long)BR; packed += k0 * BC * (long long)BR; tau_out += k0 * BC;
  double* __restrict__ rband = gband + (long long)blockIdx.x * W * BC;
  double* __restrict__ y = gy + (long long)blockIdx.x * W;
  if (b) b += k0 * BR;
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
    const int ncols_w = (k0 + k == nb_total - 1) ? last_cols : BC;   // the last slab may be narrower (fromBlockBandedPattern, SparseQRUtils.h:284)

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
      if (last ? (r < ncols_w) : (r < S)) {      // the group's last window hands over all its rows (R_g's tail triangle)
        const double val = QRK_WROW(r);
        const long long g = k * S + r;
        if (is_col) rband[g * BC + lane] = (lane >= r) ? val : 0.0;
        else if (is_rhs) y[g] = val;
      }
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
// Phase 2: chase the OV-row bulge through the group triangles.  Lane l owns column l of the current window of the band
// (lane BC the right-hand side): Bg[i] = bulge row i of that column.  Step q (= group * W + local row; the steps of all
// groups are contiguous) takes pivot row q of the group triangles: reflector on [pivot entry of lane c ; Bg[0..OV) of lane
// c], c = the row's column inside its window; the updated pivot row is a final row of R — or, for the last OV rows of a
// group, part of the bulge handed to the next group.
// ONE rolled loop over q: a single warp cannot hide instruction fetch, so the body must stay inside the instruction
// cache (the unrolled-by-column form of this loop spent 15 % of its cycles waiting for instructions).  Pivot rows stream
// through a cp.async ring D steps ahead of the chain; all stores are branch-free (pointer selects).
// craw[q*OV ..] (raw tail), cs[2q] = tau, cs[2q+1] = inv keep the reflector v = [1; inv * raw] for later applications.
// ---------------------------------------------------------------------------------------------
template <int BC, int OV>
__global__ void __launch_bounds__(32, 1)
banded_chase_kernel(const double* __restrict__ gband, const double* __restrict__ gy, double* __restrict__ rband,
                    double* __restrict__ y, double* __restrict__ craw, double* __restrict__ cs, long long nb, int last_cols,
                    int group, int have_rhs) {
  constexpr int S = BC - OV, NO = OV > 0 ? OV : 1, RING = 32, PW = BC + 2;
  static_assert(BC % 2 == 0 && BC + 1 <= 32, "16-byte cp.async of whole band rows; one lane per column plus the rhs");
  __shared__ __align__(16) double prow[RING][PW];       // pivot rows: BC entries + the right-hand side value
  __shared__ __align__(16) double sv[2][NO];            // published raw tail (double buffered by step parity)
  __shared__ __align__(16) double trash[NO * 33];       // dump slots of the non-pivot lanes, entry i of lane l at i * 33 + l
  __shared__ double hand[NO][33];                       // rows handed to the next group
  const int lane = threadIdx.x;
  const bool is_col = lane < BC, is_rhs = (lane == BC) && have_rhs;
  const long long W = (long long)(group - 1) * S + BC;
  const long long ngroups = (nb + group - 1) / group;
  const long long n_last = nb - (ngroups - 1) * group;
  const long long Q = (ngroups - 1) * W + (n_last - 1) * S + last_cols;      // steps in total

  // pivot rows are fetched 8 at a time, two batches (16 rows) ahead of the chain: one cp.async group per 8 steps
  auto prefetch8 = [&](long long q8) {                   // rows [q8, q8 + 8)
    const int slot = (int)(q8 % RING);
#pragma unroll
    for (int j = lane; j < 8 * (BC / 2); j += 32) {
      const int rr = j / (BC / 2), cc = j - rr * (BC / 2);
      if (q8 + rr < Q) cp_async16(&prow[slot + rr][2 * cc], gband + (q8 + rr) * BC + 2 * cc);
    }
    if (lane < 8 && have_rhs && q8 + lane < Q) cp_async8(&prow[slot + lane][BC], gy + q8 + lane);
    cp_async_commit();
  };
  prefetch8(0);
  prefetch8(8);
  cp_async_wait<0>();
  __syncwarp();

  double Bg[NO];
#pragma unroll
  for (int i = 0; i < NO; i++) Bg[i] = 0.0;
  double pv = (is_col || is_rhs) ? prow[0][lane] : 0.0;

  long long gi = 0;                         // group
  long long n_g = (nb < group) ? nb : group;
  double* p_raw = craw + (lane < OV ? lane : 0);
  double* p_cs = cs + (lane < 2 ? lane : 0);
  double* p_out = is_col ? rband + lane : y;                    // next final row of R (this lane's entry) / of y
  const int out_stride = is_col ? BC : 1;
  long long lw = 0;                         // window inside the group
  int c = 0;                                // column of the pivot row inside its window
  bool lastgroup = (ngroups == 1);
  for (long long q = 0; q < Q; q++) {
    const bool lastw = (lw == n_g - 1);
    // ---- publish lane c's raw tail (every lane stores: lane c to the buffer, the others to a dump slot -> no divergence)
    double* dst = (lane == c) ? sv[q & 1] : trash + lane;       // (a dump word per lane and entry: no write-write overlap for racecheck)
    const int dstep = (lane == c) ? 1 : 33;
#pragma unroll
    for (int i = 0; i < OV; i++) dst[i * dstep] = Bg[i];
    double tq[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < OV; i++) tq[i & 3] = fma(Bg[i], Bg[i], tq[i & 3]);
    const double tailSq = (tq[0] + tq[1]) + (tq[2] + tq[3]);
    double beta, inv, tau;
    householder_scalars(pv, tailSq, OV == 0, beta, inv, tau);
    if ((q & 7) == 0) {
      prefetch8(q + 16);
      cp_async_wait<1>();                    // rows < q + 16 have landed
    }
    __syncwarp();
    const double pv_next = (is_col || is_rhs) ? prow[(q + 1) % RING][lane] : 0.0;
    double t[NO];
    double dq[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < OV; i++) {
      t[i] = sv[q & 1][i];
      dq[i & 3] = fma(t[i], Bg[i], dq[i & 3]);
    }
    const double mine = (lane < OV) ? sv[q & 1][lane < OV ? lane : 0] : 0.0;
    const double dot = (dq[0] + dq[1]) + (dq[2] + dq[3]);
    const double tau_c = __shfl_sync(0xffffffffu, tau, c);
    const double inv_c = __shfl_sync(0xffffffffu, inv, c);
    const double w = (lane > c) ? tau_c * fma(inv_c, dot, pv) : 0.0;
    const double z = (lane == c) ? 1.0 : w * inv_c;            // lane c: Bg - 1 * t = 0 exactly, its column is annihilated
    pv -= w;
    if (lane == c) pv = beta;
#pragma unroll
    for (int i = 0; i < OV; i++) Bg[i] = fma(-z, t[i], Bg[i]);
    // ---- the reflector, for later applications (running per-lane pointers: no 64-bit index arithmetic in the loop)
    if (lane < OV) *p_raw = mine;
    if (lane < 2) *p_cs = lane ? inv_c : tau_c;
    p_raw += OV; p_cs += 2;
    // ---- the updated pivot row: final row of R, or handed to the next group
    const double outv = (lane >= c) ? pv : 0.0;
    if (!(lastw && !lastgroup && c >= S)) {
      if (is_col || is_rhs) *p_out = outv;
      p_out += out_stride;
    } else {
      hand[c - S][lane] = outv;
    }
    // ---- position of the next step
    pv = pv_next;
    c++;
    if (!lastw) {
      if (c == S) {                          // next window: the bulge moves S columns to the right
        c = 0; lw++;
#pragma unroll
        for (int i = 0; i < OV; i++) {
          const double shifted = __shfl_down_sync(0xffffffffu, Bg[i], S);
          Bg[i] = (lane == BC) ? Bg[i] : ((lane < BC - S) ? shifted : 0.0);
        }
      }
    } else if (!lastgroup && c == BC) {      // next group: the handed-over rows are the new bulge
      __syncwarp();
#pragma unroll
      for (int i = 0; i < OV; i++)
        Bg[i] = (lane == BC) ? hand[i][BC] : ((lane < BC - S) ? hand[i][lane + S] : 0.0);
      gi++; lw = 0; c = 0;
      n_g = (nb - gi * group < group) ? (nb - gi * group) : group;
      lastgroup = (gi == ngroups - 1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Q^T applied to a new right-hand side on a stored factorisation, phase 1 (parallel over the groups): the window sweep of
// banded_factor_kernel with only the vector; lane i owns slab row i, the OV carried entries are replicated in every lane.
// Output: gy, the group's pivot-row values (W per group), input of banded_chase_apply_kernel.
// ---------------------------------------------------------------------------------------------
template <int BR, int BC, int OV>
__global__ void __launch_bounds__(32, 1)
banded_apply_qt_kernel(const double* __restrict__ packed, const double* __restrict__ tau_in, const double* __restrict__ b,
                       double* __restrict__ gy, long long nb_total, int last_cols, int group, double* __restrict__ comp,
                       long long ldb, long long ldgy, long long ldcomp) {
  using G = BandedCfg<BR, BC, OV>;
  constexpr int S = G::S, M = G::M;
  static_assert(BR <= 32, "one lane per slab row");
  const int lane = threadIdx.x;
  // several right-hand sides: blockIdx.y selects the column (leading dimensions ldb / ldgy / ldcomp)
  b += (long long)blockIdx.y * ldb; gy += (long long)blockIdx.y * ldgy;
  if (comp) comp += (long long)blockIdx.y * ldcomp;
  const long long k0 = (long long)blockIdx.x * group;
  const long long nb = (nb_total - k0 < group) ? (nb_total - k0) : group;
  const long long W = (long long)(group - 1) * S + BC;
  packed += k0 * BC * (long long)BR; tau_in += k0 * BC; b += k0 * BR;
  double* __restrict__ y = gy + (long long)blockIdx.x * W;
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
    const int ncols_w = (k0 + k == nb_total - 1) ? last_cols : BC;
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
      if (last ? (r < ncols_w) : (r < S)) {
        const double val = (r < OV) ? cy[r < OV ? r : 0] : __shfl_sync(0xffffffffu, bi, r >= OV ? r - OV : 0);
        if (lane == 0) y[k * S + r] = val;
      }
    }
    // the annihilated rows of the window: the part of Q^T b outside range(A) (only kept when the caller wants the FULL
    // orthogonal transform, e.g. the border of a block-angular matrix whose left block is banded).  Layout of comp:
    // [nb_total * (M - BC) window rows | groups * OV bulge rows of the chase | BC - last_cols rows of the last window]
    if (comp) {
#pragma unroll
      for (int r = BC; r < M; r++) {
        const double val = (r < OV) ? cy[r < OV ? r : 0] : __shfl_sync(0xffffffffu, bi, r >= OV ? r - OV : 0);
        if (lane == 0) comp[(k0 + k) * (long long)(M - BC) + (r - BC)] = val;
      }
      if (ncols_w < BC) {                       // the matrix's last slab is narrower: its rows [last_cols, BC) carry no pivot
        const long long groups = (nb_total + group - 1) / group;
        double* tail = comp + nb_total * (long long)(M - BC) + groups * OV;
#pragma unroll
        for (int r = 0; r < BC; r++) {
          const double val = (r < OV) ? cy[r < OV ? r : 0] : __shfl_sync(0xffffffffu, bi, r >= OV ? r - OV : 0);
          if (lane == 0 && r >= ncols_w) tail[r - ncols_w] = val;
        }
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
// Q applied on a stored factorisation, phase 1 run BACKWARDS (parallel over the groups; it follows the backward chase):
// gy holds the group's pivot-row values, the annihilated rows are zero (zero complement), the windows are undone from the
// group's last to its first and every reflector of a window from the last to the first (H is symmetric).  Output: the
// slab rows, x = Q1 y.  The virtual zero rows above a group's first slab come out as ~0 and are dropped.
// ---------------------------------------------------------------------------------------------
template <int BR, int BC, int OV>
__global__ void __launch_bounds__(32, 1)
banded_apply_q_kernel(const double* __restrict__ packed, const double* __restrict__ tau_in, const double* __restrict__ gy,
                      double* __restrict__ x, long long nb_total, int last_cols, int group) {
  using G = BandedCfg<BR, BC, OV>;
  constexpr int S = G::S, NO = OV > 0 ? OV : 1;
  static_assert(BR <= 32, "one lane per slab row");
  const int lane = threadIdx.x;
  const long long k0 = (long long)blockIdx.x * group;
  const long long nb = (nb_total - k0 < group) ? (nb_total - k0) : group;
  const long long W = (long long)(group - 1) * S + BC;
  packed += k0 * BC * (long long)BR; tau_in += k0 * BC; x += k0 * BR;
  const double* __restrict__ y = gy + (long long)blockIdx.x * W;
  double cin[NO];                                  // window rows S..BC-1 as left by the window to the right
#pragma unroll
  for (int i = 0; i < NO; i++) cin[i] = 0.0;
  for (long long k = nb - 1; k >= 0; --k) {
    const bool last = (k == nb - 1);
    const int ncols_w = (k0 + k == nb_total - 1) ? last_cols : BC;
    double vv[BC], tt[BC];
#pragma unroll
    for (int c = 0; c < BC; c++) {
      vv[c] = (lane < BR) ? packed[(k * BC + c) * (long long)BR + lane] : 0.0;
      tt[c] = tau_in[k * BC + c];
    }
    // window rows: r < S pivot rows of this window; S <= r < BC from the right (or the group's tail rows); r >= BC zero
    double cy[NO];
#pragma unroll
    for (int r = 0; r < OV; r++) {
      if (r < S) cy[r] = y[k * S + r];
      else cy[r] = last ? ((r < ncols_w) ? y[k * S + r] : 0.0) : cin[r >= S ? r - S : 0];
    }
    double bi = 0.0;
    {
      const int r = OV + lane;
      if (lane < BR && r < BC) {
        if (r < S || last) bi = (r < (last ? ncols_w : S)) ? y[k * S + r] : 0.0;
        else {
#pragma unroll
          for (int j = 0; j < OV; j++) if (r - S == j) bi = cin[j];
        }
      }
    }
#pragma unroll
    for (int c = BC - 1; c >= 0; --c) {
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
    if (lane < BR) x[k * BR + lane] = bi;
#pragma unroll
    for (int i = 0; i < OV; i++) cin[i] = cy[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Q^T (forward) or Q (backward) of the chase reflectors applied to a vector of pivot-row values, phase 2 of an application
// on stored factors.  One warp, sequential over the steps q (banded_chase_kernel's order, reversed for Q); the OV bulge
// entries u live in registers (replicated in every lane), the reflectors stream through a cp.async ring.
//   forward : p = gy[q];  w = tau (p + v.u);  p -= w;  u -= w v   (v = inv * raw tail);   p -> y[global row]  (or -> next group's u for the last
//             OV rows of a group)
//   backward: the same reflector (H is symmetric) in reverse order; a group starts from u = 0 (zero complement), its
//             last OV pivot rows take the u left by the group to the right;  p -> gy[q]
// ---------------------------------------------------------------------------------------------
template <int BC, int OV, bool BACKWARD>
__global__ void __launch_bounds__(32, 1)
banded_chase_apply_kernel(const double* __restrict__ craw, const double* __restrict__ cs, const double* __restrict__ in,
                          double* __restrict__ out, long long nb, int last_cols, int group, double* __restrict__ ucomp) {
  constexpr int S = BC - OV, NO = OV > 0 ? OV : 1, CH = 32, NST = 4;
  __shared__ __align__(16) double rv[NST][CH * NO];
  __shared__ __align__(16) double rt[NST][2 * CH];     // tau, inv per step
  __shared__ double rp[NST][CH];
  __shared__ double hand[NO];                          // the OV values that cross a group boundary
  const int lane = threadIdx.x;
  const long long W = (long long)(group - 1) * S + BC;
  const long long ngroups = (nb + group - 1) / group;
  const long long n_last = nb - (ngroups - 1) * group;           // slabs of the last group
  const long long W_last = (n_last - 1) * S + last_cols;         // its pivot rows
  const long long Q = (ngroups - 1) * W + W_last;                // steps in total
  const long long nchunks = (Q + CH - 1) / CH;
  const long long GS = (long long)group * S;                     // final rows per (full) group

  // global row of step q (for the vector that is NOT indexed by q): final rows only
  auto chunk_of = [&](long long i) { return BACKWARD ? nchunks - 1 - i : i; };
  auto prefetch = [&](long long i) {
    if (i < nchunks) {
      const long long ck = chunk_of(i);
      const int st = (int)(i % NST);
      const long long q0 = ck * CH;
      const int valid = (int)((Q - q0 < CH) ? (Q - q0) : CH);
      if (OV > 0) {
        const double* src = craw + q0 * OV;
        for (int j = lane; j < valid * OV / 2; j += 32) cp_async16(&rv[st][2 * j], src + 2 * j);
        if ((valid * OV) & 1) { if (lane == 0) cp_async8(&rv[st][valid * OV - 1], src + valid * OV - 1); }
      }
      if (lane < valid) {
        const long long q = q0 + lane;
        cp_async16(&rt[st][2 * lane], cs + 2 * q);
        const long long gi = q / W, r = q - gi * W;
        const bool handed = (gi < ngroups - 1) && (r >= GS);      // pivot row that belongs to the next group's bulge
        if (BACKWARD) { if (!handed) cp_async8(&rp[st][lane], in + gi * GS + r); }
        else cp_async8(&rp[st][lane], in + q);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int i = 0; i < NST - 1; i++) prefetch(i);

  double u[NO];
#pragma unroll
  for (int i = 0; i < NO; i++) u[i] = 0.0;
  if (lane < NO) hand[lane] = 0.0;
  __syncwarp();

  for (long long i = 0; i < nchunks; i++) {
    prefetch(i + NST - 1);
    cp_async_wait<NST - 1>();
    __syncwarp();
    const int st = (int)(i % NST);
    const long long q0 = chunk_of(i) * CH;
    const int valid = (int)((Q - q0 < CH) ? (Q - q0) : CH);
    double keep = 0.0;                                   // lane j keeps the result of the chunk's step j
    long long gi = (q0 + (BACKWARD ? valid - 1 : 0)) / W;          // one division per chunk, then incremental
    long long r = q0 + (BACKWARD ? valid - 1 : 0) - gi * W;
    for (int jj = 0; jj < valid; jj++) {
      const int j = BACKWARD ? valid - 1 - jj : jj;
      const long long W_g = (gi == ngroups - 1) ? W_last : W;
      const bool handed = (gi < ngroups - 1) && (r >= GS);
      if (BACKWARD && r == W_g - 1) {                    // entering a group from its end: zero complement
#pragma unroll
        for (int k = 0; k < NO; k++) u[k] = 0.0;
      }
      double p = (BACKWARD && handed) ? hand[r - GS] : rp[st][j];
      const double* v = &rv[st][j * NO];
      double dq[4] = {0.0, 0.0, 0.0, 0.0};
      double vr[NO];
#pragma unroll
      for (int k = 0; k < OV; k++) { vr[k] = v[k]; dq[k & 3] = fma(vr[k], u[k], dq[k & 3]); }
      const double inv = rt[st][2 * j + 1];                  // v = [1; inv * raw]
      const double w = rt[st][2 * j] * fma(inv, (dq[0] + dq[1]) + (dq[2] + dq[3]), p);
      const double wz = w * inv;
      p -= w;
#pragma unroll
      for (int k = 0; k < OV; k++) u[k] = fma(-wz, vr[k], u[k]);
      if (lane == j) keep = p;
      if (!BACKWARD) {
        if (handed) { if (lane == 0) hand[r - GS] = p; }
        if (r == W_g - 1 && ucomp != nullptr && lane < OV) {   // the annihilated bulge rows of this group (forward only)
          double mine = 0.0;
#pragma unroll
          for (int k = 0; k < OV; k++) if (lane == k) mine = u[k];
          ucomp[gi * OV + lane] = mine;
        }
        if (r == W_g - 1 && gi < ngroups - 1) {          // leaving a group: its last OV pivot rows are the next bulge
          __syncwarp();
#pragma unroll
          for (int k = 0; k < OV; k++) u[k] = hand[k];
          __syncwarp();                                  // every lane has read hand[] before lane 0 writes the next group's rows
        }
      } else if (r == 0) {                               // left end of a group: u is what the group to the left handed over
        __syncwarp();
        if (lane < OV) {
          double mine = 0.0;
#pragma unroll
          for (int k = 0; k < OV; k++) if (lane == k) mine = u[k];
          hand[lane] = mine;
        }
        __syncwarp();
      }
      if (BACKWARD) { if (--r < 0) { gi--; r = W - 1; } }
      else if (++r == W_g) { gi++; r = 0; }
    }
    // coalesced store of the chunk's results
    if (lane < valid) {
      const long long q = q0 + lane;
      const long long gi = q / W, r = q - gi * W;
      const bool handed = (gi < ngroups - 1) && (r >= GS);
      if (BACKWARD) out[q] = keep;
      else if (!handed) out[gi * GS + r] = keep;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Forward chase application on SEVERAL vectors at once (the border columns of a block-angular matrix whose left block is
// banded): lane l of CTA c owns column 32 c + l — its OV bulge entries in registers, its hand-over slots in shared
// memory — and all lanes share the reflector stream.  The steps are independent across the columns, so the sequential
// sweep over the n_cols + groups OV steps is paid once per 32 columns.
//   in : gy of column j at in + j ldin;  out: thin part at out + j ldout, bulge complement at ucomp + j lducomp
// ---------------------------------------------------------------------------------------------
template <int BC, int OV>
__global__ void __launch_bounds__(32, 1)
banded_chase_apply_multi_kernel(const double* __restrict__ craw, const double* __restrict__ cs, const double* __restrict__ in,
                                long long ldin, double* __restrict__ out, long long ldout, double* __restrict__ ucomp,
                                long long lducomp, int ncols, long long nb, int last_cols, int group) {
  constexpr int S = BC - OV, NO = OV > 0 ? OV : 1, CH = 32, NST = 4;
  __shared__ __align__(16) double rv[NST][CH * NO];
  __shared__ __align__(16) double rt[NST][2 * CH];
  __shared__ double hand[NO][33];
  const int lane = threadIdx.x;
  const int col = (int)blockIdx.x * 32 + lane;
  const bool active = col < ncols;
  const double* my_in = in + (long long)(active ? col : 0) * ldin;
  double* my_out = out + (long long)(active ? col : 0) * ldout;
  double* my_ucomp = ucomp ? ucomp + (long long)(active ? col : 0) * lducomp : nullptr;
  const long long W = (long long)(group - 1) * S + BC;
  const long long ngroups = (nb + group - 1) / group;
  const long long n_last = nb - (ngroups - 1) * group;
  const long long W_last = (n_last - 1) * S + last_cols;
  const long long Q = (ngroups - 1) * W + W_last;
  const long long nchunks = (Q + CH - 1) / CH;
  const long long GS = (long long)group * S;

  auto prefetch = [&](long long i) {
    if (i < nchunks) {
      const int st = (int)(i % NST);
      const long long q0 = i * CH;
      const int valid = (int)((Q - q0 < CH) ? (Q - q0) : CH);
      if (OV > 0) {
        const double* src = craw + q0 * OV;
        for (int j = lane; j < valid * OV / 2; j += 32) cp_async16(&rv[st][2 * j], src + 2 * j);
        if ((valid * OV) & 1) { if (lane == 0) cp_async8(&rv[st][valid * OV - 1], src + valid * OV - 1); }
      }
      if (lane < valid) cp_async16(&rt[st][2 * lane], cs + 2 * (q0 + lane));
    }
    cp_async_commit();
  };
#pragma unroll
  for (int i = 0; i < NST - 1; i++) prefetch(i);

  double u[NO];
#pragma unroll
  for (int i = 0; i < NO; i++) u[i] = 0.0;
  long long gi = 0, r = 0;
  for (long long i = 0; i < nchunks; i++) {
    prefetch(i + NST - 1);
    cp_async_wait<NST - 1>();
    __syncwarp();
    const int st = (int)(i % NST);
    const long long q0 = i * CH;
    const int valid = (int)((Q - q0 < CH) ? (Q - q0) : CH);
    double pin[CH];                                     // this column's pivot-row values of the chunk (independent loads)
#pragma unroll
    for (int j = 0; j < CH; j++) pin[j] = (active && j < valid) ? my_in[q0 + j] : 0.0;
#pragma unroll
    for (int j = 0; j < CH; j++) {
      if (j < valid) {                                  // uniform
        const long long W_g = (gi == ngroups - 1) ? W_last : W;
        const bool handed = (gi < ngroups - 1) && (r >= GS);
        const double* v = &rv[st][j * NO];
        double dq[4] = {0.0, 0.0, 0.0, 0.0};
        double vr[NO];
#pragma unroll
        for (int k = 0; k < OV; k++) { vr[k] = v[k]; dq[k & 3] = fma(vr[k], u[k], dq[k & 3]); }
        const double inv = rt[st][2 * j + 1];
        double p = pin[j];
        const double w = rt[st][2 * j] * fma(inv, (dq[0] + dq[1]) + (dq[2] + dq[3]), p);
        const double wz = w * inv;
        p -= w;
#pragma unroll
        for (int k = 0; k < OV; k++) u[k] = fma(-wz, vr[k], u[k]);
        if (handed) hand[r - GS][lane] = p;             // own slot: no cross-lane traffic
        else if (active) my_out[gi * GS + r] = p;
        if (r == W_g - 1) {                              // end of the group: annihilated bulge out, handed rows in
          if (my_ucomp && active) {
#pragma unroll
            for (int k = 0; k < OV; k++) my_ucomp[gi * OV + k] = u[k];
          }
          if (gi < ngroups - 1) {
#pragma unroll
            for (int k = 0; k < OV; k++) u[k] = hand[k][lane];
          }
        }
        if (++r == W_g) { gi++; r = 0; }
      }
    }
    __syncwarp();
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
