// dense_tri_reg.cuh — Eigen's ColPivHouseholderQR on the M x (M + nrhs) triangle of a wide border (second stage of the
// block-angular path with a ColPiv right solver, BlockAngularSparseQR.h:368 / [Eigen] ColPivHouseholderQR.h), the whole matrix
// REGISTER resident in one thread-block cluster.
//
// dense_tri_colpiv_kernel (dense_blocked.cuh) keeps the triangle in the cluster's shared memory; its per-phase trace
// (tools/tri_trace.py, k = 64 of 384) shows 7.4k of the 13.5k cycles of a column step in the trailing update, which is bound
// by shared-memory bandwidth: every column entry is read twice and written once per step.  Here every warp of the cluster
// OWNS up to 7 columns for the whole factorisation, 12 rows per lane each (8 CTAs x 8 warps x 7 slots = 448 columns of up to
// 384 rows = 84 doubles per lane), so a step touches shared memory only for the reflector (3 KB) and the norm tables.
// Nothing is ever swapped: a pivot exchange only exchanges the LOGICAL positions of two physical columns (a replicated
// position table), the data stays where it is and is written to its final position at the end.
//   step k:  (1) local pivot candidates (first maximum of the downdated norms, lowest logical index on ties), pushed to
//                every CTA;  cluster barrier
//            (2) global first maximum, logical positions exchanged in every CTA's tables
//            (3) the warp that owns the pivot column builds the reflector from its registers; its CTA pushes v, tau, beta
//                into every CTA's shared memory;  cluster barrier
//            (4) every warp updates its active columns in registers (two groups of slots, each group's chains side by
//                side), LAWN-176 norm downdates one lane per column
// Limits: M <= 384, M + nrhs <= 448 (reference tests 4 / 6: 384 + 1); larger or tiny problems keep the shared-memory kernel.
#pragma once
#include "dense_blocked.cuh"

namespace qrk {

constexpr int kTrThreads = 256, kTrWarps = kTrThreads / 32, kTrSlots = 7, kTrRows = 12;
constexpr int kTrLocal = kTrWarps * kTrSlots;                   // physical columns per CTA (56)
constexpr int kTrMaxM = 32 * kTrRows;                           // 384
constexpr int kTrMaxCols = kDbCluster * kTrLocal;               // 448

__host__ __device__ inline bool tri_reg_fits(int M, int nrhs) { return M > 64 && M <= kTrMaxM && M + nrhs <= kTrMaxCols; }

// (4) for the slots [s0, s0 + cnt) of one warp: their chains run side by side.  (Reduce-scatter reductions -- 6 64-bit
// shuffles for four values instead of 20 -- were measured SLOWER here: 6.6k against 5.5k cycles for this phase.)
template <int s0, int cnt>
__device__ __forceinline__ void tri_reg_update(double (&x)[kTrSlots][kTrRows], const double* vloc, const int* logj, double* upd,
                                               double* dir, double tau, int k, int M, int NC, int rank, int warp, int lane) {
  const int krs = k >> 5, klane = k & 31;
  bool act[cnt];
  int lj[cnt];
  double dot[cnt], nsq[cnt], ak[cnt];
#pragma unroll
  for (int q = 0; q < cnt; q++) {
    const int l = warp + kTrWarps * (s0 + q), g = rank + kDbCluster * l;
    lj[q] = logj[l];
    act[q] = g < NC && lj[q] > k;
    dot[q] = 0.0; nsq[q] = 0.0; ak[q] = 0.0;
  }
  // Row slots above the pivot row's slot hold finished rows of R (the reflector is zero there): they are skipped by
  // warp-uniform branches on the slot index, so nothing in the loops is a per-element select (the first version of this
  // kernel spent 1755 of its 4096 instructions on FSEL) and the work shrinks as k advances.
#pragma unroll
  for (int rs = 0; rs < kTrRows; rs++) {
    if (rs >= krs) {
      const double v = vloc[rs * 32 + lane];
#pragma unroll
      for (int q = 0; q < cnt; q++) dot[q] = fma(v, x[s0 + q][rs], dot[q]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int q = 0; q < cnt; q++) dot[q] += __shfl_xor_sync(0xffffffffu, dot[q], o);
  }
#pragma unroll
  for (int q = 0; q < cnt; q++) dot[q] = act[q] ? tau * dot[q] : 0.0;      // an inactive column is "updated" by zero: unchanged
#pragma unroll
  for (int rs = 0; rs < kTrRows; rs++) {
    if (rs > krs) {
      const double v = vloc[rs * 32 + lane];
#pragma unroll
      for (int q = 0; q < cnt; q++) {
        const double y = fma(-v, dot[q], x[s0 + q][rs]);
        x[s0 + q][rs] = y;
        nsq[q] = fma(y, y, nsq[q]);
      }
    } else if (rs == krs) {
      const double v = vloc[rs * 32 + lane];
      const bool below = lane > klane;
#pragma unroll
      for (int q = 0; q < cnt; q++) {
        const double y = fma(-v, dot[q], x[s0 + q][rs]);
        x[s0 + q][rs] = y;
        ak[q] = y;
        if (below) nsq[q] = fma(y, y, nsq[q]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int q = 0; q < cnt; q++) nsq[q] += __shfl_xor_sync(0xffffffffu, nsq[q], o);
  }
  // lane q finishes column q of the group (Eigen's LAWN-176 rule; correctly rounded fast reciprocal / square root)
  double my_ak = 0.0, my_nsq = 0.0;
  int my_l = 0, my_j = M;
  bool my_act = false;
#pragma unroll
  for (int q = 0; q < cnt; q++) {
    const double a0 = __shfl_sync(0xffffffffu, ak[q], klane);
    if (lane == q) { my_ak = a0; my_nsq = nsq[q]; my_l = warp + kTrWarps * (s0 + q); my_j = lj[q]; my_act = act[q]; }
  }
  if (lane < cnt && my_act && my_j < M) {
    const double u = upd[my_l];
    if (u != 0.0) {
      double t = fabs(my_ak) * fast_rcp(u);
      t = (1.0 + t) * (1.0 - t);
      t = t < 0.0 ? 0.0 : t;
      const double qd = u * fast_rcp(dir[my_l]);
      const double t2 = t * (qd * qd);
      if (t2 <= 1.4901161193847656e-08) {                    // sqrt(eps): recompute
        double nrm;
        (void)fast_rsqrt(my_nsq, nrm);
        dir[my_l] = nrm; upd[my_l] = nrm;
      } else {
        double st;
        (void)fast_rsqrt(t, st);
        upd[my_l] = u * st;
      }
    }
  }
}

__global__ void __cluster_dims__(kDbCluster, 1, 1) __launch_bounds__(kTrThreads) dense_tri_colpiv_reg_kernel(DenseBorder d) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ __align__(16) double vloc[kTrMaxM];    // this step's reflector: 0 above row k, 1 at row k, essential part below
  __shared__ double hdr[2];                          // tau, beta of the published reflector
  __shared__ double upd[kTrLocal], dir[kTrLocal];    // m_colNormsUpdated / m_colNormsDirect of the local physical columns
  __shared__ int logj[kTrLocal];                     // current logical position of the local physical columns
  __shared__ int where[kTrMaxCols];                  // logical position -> physical column (replicated in every CTA)
  __shared__ double candv[2][kDbCluster];            // pivot candidates of the 8 CTAs: norm, logical position, physical column
  __shared__ int candj[2][kDbCluster], candg[2][kDbCluster];
  __shared__ double sred[kTrWarps];
  __shared__ double gmaxs[kDbCluster];
  const int M = d.M, NC = M + d.nrhs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster.block_rank();

  // ---- load: slot s of this warp is local column l = warp + 8 s, physical (= original) column g = rank + 8 l
  double x[kTrSlots][kTrRows];
#pragma unroll
  for (int s = 0; s < kTrSlots; s++) {
    const int g = rank + kDbCluster * (warp + kTrWarps * s);
#pragma unroll
    for (int rs = 0; rs < kTrRows; rs++) {
      const int i = rs * 32 + lane;
      x[s][rs] = (g < NC && i < M) ? d.A[(size_t)g * M + i] : 0.0;
    }
  }
  double lmax = 0.0;
#pragma unroll
  for (int s = 0; s < kTrSlots; s++) {
    const int l = warp + kTrWarps * s, g = rank + kDbCluster * l;
    double q = 0.0;
#pragma unroll
    for (int rs = 0; rs < kTrRows; rs++) q = fma(x[s][rs], x[s][rs], q);
    q = warp_sum(q);
    const double nrm = sqrt(q);
    if (lane == 0) { upd[l] = nrm; dir[l] = nrm; logj[l] = g; }
    if (g < M) lmax = fmax(lmax, nrm);
  }
  for (int j = tid; j < kTrMaxCols; j += kTrThreads) where[j] = j;
  if (lane == 0) sred[warp] = lmax;
  cluster.sync();                                      // every CTA of the cluster has started: its shared memory may be written
  if (tid < kDbCluster) {                              // this CTA's maximum norm into every CTA
    double m = 0.0;
    for (int w = 0; w < kTrWarps; w++) m = fmax(m, sred[w]);
    cluster.map_shared_rank(&gmaxs[0], tid)[rank] = m;
  }
  cluster.sync();
  double gmax = 0.0;
#pragma unroll
  for (int rk = 0; rk < kDbCluster; rk++) gmax = fmax(gmax, gmaxs[rk]);
  const double me = gmax * DBL_EPSILON;
  const double helper = me * me / (double)d.Nrule;               // threshold_helper
  const int size = (int)(d.Nrule < M ? d.Nrule : M);
  const int steps = size < M ? size : M;
  int nonzero_pivots = size;
  double maxpivot = 0.0;

  for (int k = 0; k < steps; k++) {
    const int par = k & 1;
    QRK_TRI_CLK(0);
    // ---- (1) local pivot candidate over the local columns whose logical position is in [k, M)
    if (warp == 0) {
      double bv = -1.0;
      int bj = 0x7fffffff, bg = 0;
      for (int l = lane; l < kTrLocal; l += 32) {
        const int j = logj[l], g = rank + kDbCluster * l;
        if (g < NC && j >= k && j < M) {
          const double u = upd[l];
          if (u > bv || (u == bv && j < bj)) { bv = u; bj = j; bg = g; }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        const int og = __shfl_xor_sync(0xffffffffu, bg, o);
        if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; bg = og; }
      }
      if (lane < kDbCluster) {
        cluster.map_shared_rank(&candv[0][0], lane)[par * kDbCluster + rank] = bv;
        cluster.map_shared_rank(&candj[0][0], lane)[par * kDbCluster + rank] = bj;
        cluster.map_shared_rank(&candg[0][0], lane)[par * kDbCluster + rank] = bg;
      }
    }
    QRK_TRI_CLK(1);
    cluster.sync();
    QRK_TRI_CLK(2);
    // ---- (2) global first maximum; the pivot column takes logical position k, the column that held k takes the pivot's
    int big = 0x7fffffff, bigg = 0;
    double bigv = -1.0;
#pragma unroll
    for (int rk = 0; rk < kDbCluster; rk++) {
      const double ov = candv[par][rk];
      const int oj = candj[par][rk];
      if (ov > bigv || (ov == bigv && oj < big)) { bigv = ov; big = oj; bigg = candg[par][rk]; }
    }
    if (nonzero_pivots == size && bigv * bigv < helper * (double)(d.Nrule - k)) nonzero_pivots = k;
    const int physk = where[k];                        // physical column at logical position k before the exchange
    __syncthreads();                                   // every thread has read where[k] and the candidates' tables
    if (tid == 0) {
      where[big] = physk; where[k] = bigg;
      if (big != k && (physk & (kDbCluster - 1)) == rank) logj[physk / kDbCluster] = big;
      if ((bigg & (kDbCluster - 1)) == rank) logj[bigg / kDbCluster] = k;
    }
    QRK_TRI_CLK(3);
    // ---- (3) the owner of the pivot column: reflector from registers, published to every CTA
    const int orank = bigg & (kDbCluster - 1), ol = bigg / kDbCluster;
    if (rank == orank) {
      if (warp == (ol & (kTrWarps - 1))) {
        const int os = ol / kTrWarps;
        const int krs = k >> 5, klane = k & 31;
#pragma unroll
        for (int s = 0; s < kTrSlots; s++) {
          if (s == os) {
            double tailSq = 0.0, ck = 0.0;
#pragma unroll
            for (int rs = 0; rs < kTrRows; rs++) {
              if (rs > krs) tailSq = fma(x[s][rs], x[s][rs], tailSq);
              else if (rs == krs) { ck = x[s][rs]; if (lane > klane) tailSq = fma(ck, ck, tailSq); }
            }
            tailSq = warp_sum(tailSq);
            const double c0 = __shfl_sync(0xffffffffu, ck, klane);
            double beta, inv, tau;
            householder_scalars(c0, tailSq, false, beta, inv, tau);      // Eigen's makeHouseholder
#pragma unroll
            for (int rs = 0; rs < kTrRows; rs++) {
              const int i = rs * 32 + lane;
              double v = 0.0;
              if (rs > krs) { v = x[s][rs] * inv; x[s][rs] = v; }
              else if (rs == krs) {
                if (lane > klane) { v = x[s][rs] * inv; x[s][rs] = v; }
                else if (lane == klane) { v = 1.0; x[s][rs] = beta; }
              }
              vloc[i] = v;
            }
            if (lane == 0) { hdr[0] = tau; hdr[1] = beta; d.tau[k] = tau; }
          }
        }
      }
      __syncthreads();
      for (int i = tid; i < kTrMaxM; i += kTrThreads) {
        const double v = vloc[i];
#pragma unroll
        for (int rk = 0; rk < kDbCluster; rk++)
          if (rk != rank) cluster.map_shared_rank(&vloc[0], rk)[i] = v;
      }
      if (tid < kDbCluster && tid != rank) {
        double* rh = cluster.map_shared_rank(&hdr[0], tid);
        rh[0] = hdr[0]; rh[1] = hdr[1];
      }
    }
    QRK_TRI_CLK(4);
    cluster.sync();                                    // the reflector and the updated tables are visible everywhere
    QRK_TRI_CLK(5);
    const double tau = hdr[0], beta = hdr[1];
    if (fabs(beta) > maxpivot) maxpivot = fabs(beta);
    // ---- (4) H_k on the active local columns (logical position > k; right-hand sides always), norms downdated
#ifdef QRK_TRI_TWO_GROUPS
    tri_reg_update<0, 4>(x, vloc, logj, upd, dir, tau, k, M, NC, rank, warp, lane);
    tri_reg_update<4, 3>(x, vloc, logj, upd, dir, tau, k, M, NC, rank, warp, lane);
#else
    tri_reg_update<0, kTrSlots>(x, vloc, logj, upd, dir, tau, k, M, NC, rank, warp, lane);      // all seven chains side by side
#endif
    QRK_TRI_CLK(6);
    __syncthreads();
    QRK_TRI_CLK(7);
  }

  // ---- write back: every physical column to its final (logical) position, P2, rank bookkeeping
#pragma unroll
  for (int s = 0; s < kTrSlots; s++) {
    const int l = warp + kTrWarps * s, g = rank + kDbCluster * l;
    if (g < NC) {
      const int j = logj[l];
#pragma unroll
      for (int rs = 0; rs < kTrRows; rs++) {
        const int i = rs * 32 + lane;
        if (i < M) d.A[(size_t)j * M + i] = x[s][rs];
      }
      if (j < M && lane == 0) d.perm[j] = g;
    }
  }
  if (rank == 0 && tid == 0) { d.scal[0] = helper; d.scal[1] = maxpivot; d.iscal[0] = nonzero_pivots; }
  cluster.sync();                                      // shared memory stays alive while peers may still write into it
}

}  // namespace qrk
