// dense_border.cuh — the dense right block of BlockAngularSparseQR for borders wider than the in-SM TSQR path
// (m2 > 8; reference test 4 uses 384 columns, test/test-qrkit.cpp:388-391).
//
// Reference path: BlockAngularSparseQR::solveRightBlock (src/QRKit/BlockAngularSparseQR.h:361-369) hands the residual
// rows J2' = (Q1^T J2)[m1:n, :] to RightSolver::compute, RightSolver = ColPivHouseholderQR<MatrixXd> in the reference's
// tests and example (test/test-qrkit.cpp:46-48, examples/ellipse_fitting.cpp:35).  This file restates that solver on
// the device, column by column, with Eigen's decision rules (first maximum of the LAWN-176 downdated norms, recompute
// test, nonzero-pivot threshold, rank()) so that P2, R2 and rank match the reference, not only x:
//   dense_piv_kernel  (1 CTA)            pivot search, column swap, reflector of column k
//   dense_upd_kernel  (1 CTA per column) H_k applied to the columns to the right and to the right-hand side, norm downdate
// The right-hand side rides along as an unpivoted extra column.  The matrix lives in global memory (L2-resident at the
// reference's sizes); this is the unblocked BLAS-2 form — the blocked compact-WY / DMMA update of bd_wy.cuh is the
// next step for this path.
#pragma once
#include "bd_generic.cuh"

namespace qrk {

struct DenseBorder {
  double* A;         // N x (M + nrhs) column-major
  long long ld, N;
  long long Nrule;   // rows of the matrix Eigen's rules refer to (= N, or the tall residual's rows when A is its M x M triangle)
  int M, nrhs;
  int pivot;         // 1: ColPivHouseholderQR (Eigen's pivot rule), 0: HouseholderQR / BlockedThinDenseQR (no pivoting),
                     // 2: BlockedThinSparseQR — ColPiv inside the panel [.., pend) only, per-panel nonzero-pivot rule (iscal[1])
  int pend = 0;      // pivot == 2: end (exclusive) of the current panel's column positions
  double *upd, *dir, *tau;
  int* perm;         // P2: perm[c] = original border column at position c
  double* scal;      // [0] threshold_helper, [1] maxpivot
  int* iscal;        // [0] nonzero_pivots
};

template <int TPB>
__device__ __forceinline__ double block_sum(double v, double* sred) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                       // sred may still be read from a previous call
  if (lane == 0) sred[warp] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < TPB / 32; w++) s += sred[w];
  return s;
}

// column norms: m_colNormsUpdated = m_colNormsDirect = ||A[:, j]||   (Eigen ColPivHouseholderQR::computeInPlace)
__global__ void __launch_bounds__(256) dense_norms_kernel(DenseBorder d) {
  __shared__ double sred[8];
  const double* col = d.A + (long long)blockIdx.x * d.ld;
  double s = 0.0;
  for (long long i = threadIdx.x; i < d.N; i += 256) s = fma(col[i], col[i], s);
  s = block_sum<256>(s, sred);
  if (threadIdx.x == 0) { const double nrm = sqrt(s); d.upd[blockIdx.x] = nrm; d.dir[blockIdx.x] = nrm; }
}

// threshold_helper = (max norm * eps)^2 / rows, nonzero_pivots = min(rows, cols), maxpivot = 0, P2 = identity
__global__ void __launch_bounds__(256) dense_prep_kernel(DenseBorder d) {
  __shared__ double sred[8];
  double m = 0.0;
  for (int j = threadIdx.x; j < d.M; j += 256) { m = fmax(m, d.upd[j]); d.perm[j] = j; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) m = fmax(m, sred[w]);
    const double me = m * DBL_EPSILON;
    d.scal[0] = me * me / (double)d.Nrule;
    d.scal[1] = 0.0;
    d.iscal[0] = (int)(d.Nrule < d.M ? d.Nrule : d.M);
  }
}

// BlockedThinSparseQR (BlockedThinSparseQR.h:238-246): start of the panel at positions [k0, k0 + pc) — the dense block Ji spans
// rows [k0, N), so the panel's ColPivHouseholderQR sees the column norms over those rows, its threshold_helper
// (max norm * eps)^2 / rows and its own nonzero-pivot count (kept as a position in iscal[1]; k0 + pc = "all nonzero so far")
__global__ void __launch_bounds__(256) thin_panel_prep_kernel(DenseBorder d, int k0, int pc) {
  __shared__ double sred[8];
  __shared__ double smax;
  if (threadIdx.x == 0) smax = 0.0;
  for (int c = 0; c < pc; c++) {
    const double* col = d.A + (long long)(k0 + c) * d.ld;
    double s = 0.0;
    for (long long i = k0 + threadIdx.x; i < d.N; i += 256) s = fma(col[i], col[i], s);
    s = block_sum<256>(s, sred);
    if (threadIdx.x == 0) { const double nrm = sqrt(s); d.upd[k0 + c] = nrm; d.dir[k0 + c] = nrm; smax = fmax(smax, nrm); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double me = smax * DBL_EPSILON;
    d.scal[0] = me * me / (double)(d.N - k0);
    d.iscal[1] = k0 + pc;
  }
}
// P2 = identity, maxpivot = 0, nonzero pivots = 0 (before the first panel)
__global__ void thin_init_kernel(DenseBorder d) {
  for (int j = threadIdx.x; j < d.M; j += blockDim.x) { d.perm[j] = j; d.tau[j] = 0.0; }
  if (threadIdx.x == 0) { d.scal[0] = 0.0; d.scal[1] = 0.0; d.iscal[0] = 0; d.iscal[1] = 0; }
}
// a deferred (zero-pivot) column at position c whose own reflector step was `step`: below that row the transformed column is
// zero (H x = beta e1); the stored reflector is dropped so that later panels see data, not a Householder vector
__global__ void thin_clear_below_kernel(DenseBorder d, int c, int step) {
  double* col = d.A + (long long)c * d.ld;
  for (long long i = step + 1 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < d.N; i += (long long)gridDim.x * blockDim.x) col[i] = 0.0;
}

// step k, part 1: first maximum of upd[k..M), swap, reflector of column k (Eigen makeHouseholderInPlace)
template <int TPB>
__global__ void __launch_bounds__(TPB) dense_piv_kernel(DenseBorder d, int k) {
  __shared__ double sred[TPB / 32];
  __shared__ double sval[TPB / 32];
  __shared__ int sidx[TPB / 32];
  __shared__ int s_big;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double bv = -1.0;
  int bj = 0x7fffffff;
  const int jend = (d.pivot == 2) ? d.pend : d.M;
  if (!d.pivot) { bv = 0.0; bj = k; }
  else for (int j = k + tid; j < jend; j += TPB) {
    const double u = d.upd[j];
    if (u > bv) { bv = u; bj = j; }                    // strict '>': the first maximum wins inside a thread
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
  }
  if (lane == 0) { sval[warp] = bv; sidx[warp] = bj; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < TPB / 32; w++)
      if (sval[w] > bv || (sval[w] == bv && sidx[w] < bj)) { bv = sval[w]; bj = sidx[w]; }
    const int size = (int)(d.Nrule < d.M ? d.Nrule : d.M);
    if (d.pivot == 2) {                                  // the panel's own ColPivHouseholderQR: m_nonzero_pivots, as a position
      if (d.iscal[1] == jend && bv * bv < d.scal[0] * (double)(d.N - k)) d.iscal[1] = k;
    } else if (d.pivot && d.iscal[0] == size && bv * bv < d.scal[0] * (double)(d.Nrule - k)) d.iscal[0] = k;
    if (bj != k) {
      double t = d.upd[k]; d.upd[k] = d.upd[bj]; d.upd[bj] = t;
      t = d.dir[k]; d.dir[k] = d.dir[bj]; d.dir[bj] = t;
      const int p = d.perm[k]; d.perm[k] = d.perm[bj]; d.perm[bj] = p;
    }
    s_big = bj;
  }
  __syncthreads();
  const int big = s_big;
  double* ck = d.A + (long long)k * d.ld;
  if (big != k) {                                        // m_qr.col(k).swap(m_qr.col(biggest_col_index)): all rows
    double* cb = d.A + (long long)big * d.ld;
    for (long long i = tid; i < d.N; i += TPB) { const double t = ck[i]; ck[i] = cb[i]; cb[i] = t; }
  }
  __syncthreads();
  double tailSq = 0.0;
  for (long long i = k + 1 + tid; i < d.N; i += TPB) tailSq = fma(ck[i], ck[i], tailSq);
  tailSq = block_sum<TPB>(tailSq, sred);
  const double c0 = ck[k];
  double beta, tau, inv;
  if (tailSq <= DBL_MIN) { tau = 0.0; beta = c0; inv = 0.0; }
  else {
    beta = sqrt(fma(c0, c0, tailSq));
    if (c0 >= 0.0) beta = -beta;
    inv = 1.0 / (c0 - beta);
    tau = (beta - c0) / beta;
  }
  __syncthreads();                                       // every thread has read ck[k]
  for (long long i = k + 1 + tid; i < d.N; i += TPB) ck[i] *= inv;
  if (tid == 0) {
    ck[k] = beta;
    d.tau[k] = tau;
    if (fabs(beta) > d.scal[1]) d.scal[1] = fabs(beta);  // m_maxpivot
  }
}

// step k, part 2: column j = k+1+blockIdx.x (or a right-hand side) <- H_k column; LAWN-176 downdate of its norm
template <int TPB>
__global__ void __launch_bounds__(TPB) dense_upd_kernel(DenseBorder d, int k) {
  __shared__ double sred[TPB / 32];
  const int tid = threadIdx.x;
  const int j = k + 1 + blockIdx.x;
  const double* v = d.A + (long long)k * d.ld;
  double* cj = d.A + (long long)j * d.ld;
  const double tau = d.tau[k];
  double dot = 0.0;
  for (long long i = k + 1 + tid; i < d.N; i += TPB) dot = fma(v[i], cj[i], dot);
  dot = block_sum<TPB>(dot, sred) + cj[k];
  const double w = tau * dot;
  double nsq = 0.0;
  for (long long i = k + 1 + tid; i < d.N; i += TPB) {
    const double a = fma(-v[i], w, cj[i]);
    cj[i] = a;
    nsq = fma(a, a, nsq);
  }
  nsq = block_sum<TPB>(nsq, sred);
  if (tid == 0) {
    const double akj = cj[k] - w;
    cj[k] = akj;
    if (j < d.M) {
      const double upd = d.upd[j];
      if (upd != 0.0) {
        double t = fabs(akj) / upd;
        t = (1.0 + t) * (1.0 - t);
        t = t < 0.0 ? 0.0 : t;
        const double q = upd / d.dir[j];
        const double t2 = t * (q * q);
        if (t2 <= 1.4901161193847656e-08) { const double nrm = sqrt(nsq); d.dir[j] = nrm; d.upd[j] = nrm; }   // sqrt(eps)
        else d.upd[j] = upd * sqrt(t);
      }
    }
  }
}

// v <- Q2^T v for a stored factorisation (solve(b) after compute()): H_{size-1} ... H_0 v, one CTA
template <int TPB>
__global__ void __launch_bounds__(TPB) dense_apply_qt_kernel(DenseBorder d, double* vec) {
  __shared__ double sred[TPB / 32];
  const int tid = threadIdx.x;
  const int size = (int)(d.N < d.M ? d.N : d.M);
  for (int k = 0; k < size; k++) {
    const double* v = d.A + (long long)k * d.ld;
    double dot = 0.0;
    for (long long i = k + 1 + tid; i < d.N; i += TPB) dot = fma(v[i], vec[i], dot);
    dot = block_sum<TPB>(dot, sred) + vec[k];
    const double w = d.tau[k] * dot;
    __syncthreads();
    for (long long i = k + 1 + tid; i < d.N; i += TPB) vec[i] = fma(-v[i], w, vec[i]);
    if (tid == 0) vec[k] -= w;
    __syncthreads();
  }
}

// v <- Q2 v: H_0 ... H_{size-1} v (matrixQ() * v of the right solver, BlockAngularSparseQR.h:568-572), one CTA
template <int TPB>
__global__ void __launch_bounds__(TPB) dense_apply_q_kernel(DenseBorder d, double* vec) {
  __shared__ double sred[TPB / 32];
  const int tid = threadIdx.x;
  const int size = (int)(d.N < d.M ? d.N : d.M);
  for (int k = size - 1; k >= 0; k--) {
    const double* v = d.A + (long long)k * d.ld;
    double dot = 0.0;
    for (long long i = k + 1 + tid; i < d.N; i += TPB) dot = fma(v[i], vec[i], dot);
    dot = block_sum<TPB>(dot, sred) + vec[k];
    const double w = d.tau[k] * dot;
    __syncthreads();
    for (long long i = k + 1 + tid; i < d.N; i += TPB) vec[i] = fma(-v[i], w, vec[i]);
    if (tid == 0) vec[k] -= w;
    __syncthreads();
  }
}

// Row signs that make a second Householder factorisation of the same matrix agree with a stored R2: sign[k] = -1 where the
// diagonals differ in sign (R is unique up to row signs; Q D with D = diag(sign) then satisfies (Q D)(D R) = A P)
__global__ void dense_diag_sign_kernel(const double* __restrict__ A, long long ld, const double* __restrict__ R2, int M, int size,
                                       double* __restrict__ sign) {
  for (int k = threadIdx.x; k < M; k += blockDim.x)
    sign[k] = (k < size && A[(long long)k * ld + k] * R2[(long long)k * M + k] < 0.0) ? -1.0 : 1.0;
}
__global__ void dense_scale_head_kernel(double* __restrict__ vec, const double* __restrict__ sign, int M) {
  for (int k = threadIdx.x; k < M; k += blockDim.x) vec[k] *= sign[k];
}

// rank (Eigen ColPivHouseholderQR::rank(): |R_ii| > |maxpivot| eps diagonalSize among the nonzero pivots), the root
// record [R2 | z2 | y2 | x2] / [P2 | rank2] shared with the TSQR path, colsPermutation()[m1 + c] = m1 + P2[c]
// (BlockAngularSparseQR.h:498-503), y2 = R2[0:rank,0:rank]^-1 z2[0:rank] (:211-217), x2 = P2 y2.
template <int TPB>
__global__ void __launch_bounds__(TPB) dense_finish_kernel(DenseBorder d, const double* z, double* root, int* root_i, int* perm_tail,
                                                           int m1, double* x2_out) {
  extern __shared__ double sy[];       // M doubles
  __shared__ int s_rank;
  const int tid = threadIdx.x, M = d.M;
  const int size = (int)(d.Nrule < M ? d.Nrule : M);
  {                                      // rank: |R_ii| > |maxpivot| eps size among the nonzero pivots, counted in parallel
    const double thresh = fabs(d.scal[1]) * (DBL_EPSILON * (double)size);
    int cnt = 0;
    if (d.pivot == 1) for (int i = tid; i < d.iscal[0]; i += TPB) cnt += (fabs(d.A[(long long)i * d.ld + i]) > thresh) ? 1 : 0;
    if (tid == 0) s_rank = 0;
    __syncthreads();
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((tid & 31) == 0 && cnt) atomicAdd(&s_rank, cnt);
    __syncthreads();
    if (tid == 0) {
      if (!d.pivot) s_rank = M;            // BlockedThinDenseQR: m_nonzeroPivots = m_R.cols() (BlockedThinDenseQR.h:132)
      if (d.pivot == 2) s_rank = d.iscal[0];   // BlockedThinSparseQR: rank() = m_nonzeroPivots, summed over the panels (:281)
      root_i[M] = s_rank;
    }
  }
  for (long long e = tid; e < (long long)M * M; e += TPB) {
    const int c = (int)(e / M), r = (int)(e - (long long)c * M);
    root[e] = (r <= c && r < d.N) ? d.A[(long long)c * d.ld + r] : 0.0;
  }
  for (int c = tid; c < M; c += TPB) { root_i[c] = d.perm[c]; perm_tail[c] = m1 + d.perm[c]; }
  __syncthreads();
  if (!z) return;
  const int rank = s_rank;
  double* rz = root + (long long)M * M;
  double* ry = rz + M;
  double* rx = ry + M;
  for (int i = tid; i < M; i += TPB) { const double zi = (i < d.N) ? z[i] : 0.0; sy[i] = zi; rz[i] = zi; }
  __syncthreads();
  // back substitution in blocks of 32 columns from the bottom: warp 0 solves the 32 x 32 diagonal block in registers
  // (lane i owns row j0 + i; one broadcast shuffle + one FMA per column), then every thread updates one row above it
  for (int j0 = ((rank - 1) / 32) * 32; j0 >= 0; j0 -= 32) {
    const int nbk = (rank - j0 < 32) ? (rank - j0) : 32;
    if (tid < 32) {
      const int lane = tid;
      double rb[32];
#pragma unroll
      for (int c = 0; c < 32; c++) rb[c] = (lane < nbk && c < nbk && c > lane) ? d.A[(long long)(j0 + c) * d.ld + j0 + lane] : 0.0;
      const double dg = (lane < nbk) ? d.A[(long long)(j0 + lane) * d.ld + j0 + lane] : 1.0;
      double sv = (lane < nbk) ? sy[j0 + lane] : 0.0;
#pragma unroll
      for (int c = 31; c >= 0; --c) {
        const double yc = __shfl_sync(0xffffffffu, sv / dg, c);      // row c is complete when the loop reaches it
        sv = (lane == c) ? yc : fma(-rb[c], yc, sv);                 // rb[c] = 0 for lanes >= c
      }
      if (lane < nbk) sy[j0 + lane] = sv;
    }
    __syncthreads();
    for (int i = tid; i < j0; i += TPB) {
      double acc = sy[i];
#pragma unroll 8
      for (int c = 0; c < nbk; c++) acc = fma(-d.A[(long long)(j0 + c) * d.ld + i], sy[j0 + c], acc);
      sy[i] = acc;
    }
    __syncthreads();
  }
  for (int c = tid; c < M; c += TPB) {
    const double yc = (c < rank) ? sy[c] : 0.0;
    ry[c] = yc;
    rx[d.perm[c]] = yc;
    if (x2_out) x2_out[d.perm[c]] = yc;
  }
}

// ytop[i] -= sum_c Atop[i, c] x2[c]   (the border columns of R times y2, BlockAngularSparseQR.h:296-301 / :211)
__global__ void __launch_bounds__(256) dense_top_kernel(const double* __restrict__ atop, long long ld, long long m1, int M,
                                                        const double* __restrict__ x2, double* __restrict__ ytop) {
  extern __shared__ double sx[];
  for (int c = threadIdx.x; c < M; c += 256) sx[c] = x2[c];
  __syncthreads();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= m1) return;
  double s0 = 0.0, s1 = 0.0;
  int c = 0;
  for (; c + 1 < M; c += 2) {
    s0 = fma(atop[(long long)c * ld + i], sx[c], s0);
    s1 = fma(atop[(long long)(c + 1) * ld + i], sx[c + 1], s1);
  }
  if (c < M) s0 = fma(atop[(long long)c * ld + i], sx[c], s0);
  ytop[i] -= s0 + s1;
}

// x1 = P1 R1^-1 ytop for the block-diagonal left factor: one warp per diagonal block, column-oriented back substitution
template <int WPC>
__global__ void __launch_bounds__(32 * WPC) bd_rsolve_kernel(BlockIndex bi, long long nb, const double* __restrict__ packed,
                                                            const int* __restrict__ perm, const double* __restrict__ y,
                                                            double* __restrict__ x, int max_c) {
  extern __shared__ double smem_rs[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long blk = (long long)blockIdx.x * WPC + warp;
  if (blk >= nb) return;
  int r, c;
  long long vo, ro, co;
  bi.get(blk, r, c, vo, ro, co);
  double* v = smem_rs + (size_t)warp * max_c;
  const double* P = packed + vo;
  for (int i = lane; i < c; i += 32) v[i] = y[co + i];
  __syncwarp();
  for (int j = c - 1; j >= 0; --j) {
    const double yj = v[j] / P[(size_t)j * r + j];
    __syncwarp();
    if (lane == 0) v[j] = yj;
    for (int i = lane; i < j; i += 32) v[i] = fma(-P[(size_t)j * r + i], yj, v[i]);
    __syncwarp();
  }
  for (int j = lane; j < c; j += 32) x[perm ? perm[co + j] : co + j] = v[j];
}

}  // namespace qrk
