// dense_blocked.cuh — blocked (compact-WY) unpivoted Householder QR of the tall dense border residual, trailing update on
// the FP64 tensor cores (mma.sync m8n8k4 f64 = SASS DMMA.8x8x4).
//
// Reference path: BlockedThinDenseQR (src/QRKit/BlockedThinDenseQR.h:104-176) factors the dense right block of
// BlockAngularSparseQR panel by panel (panel width SuggestedBlockCols), builds Y / T per panel
// (BlockedThinQRBase::computeBlockedRepresentation) and applies  A[:,j] += Y (T^T (Y^T A[:,j]))  ONE trailing column at a
// time as GEMVs (BlockedThinQRBase.h:308-333, updateMat).  Here a panel is 8 columns and the same update is two small
// GEMMs per 8-column block of the trailing matrix:  W = V^T A2  (8 x 8, contraction over the rows) and  A2 -= V (T^T W),
// both issued as DMMA.  The right-hand side rides along as one more trailing column.
//
//   dense_panel_kernel<RPT>   ONE thread-block cluster of 8 CTAs x 512 threads, rows across threads (RPT rows each, the panel
//                             in registers): 8 column steps, one fused cluster-wide reduction per step through distributed
//                             shared memory (the column's squared tail norm and its dot products with the columns to its
//                             right), then V^T V in one more reduction and the 8 x 8 triangular factor T
//                             (H_0 ... H_7 = I - V T V^T)
//   dense_wy_w_kernel /       the trailing update as two launches over (8-column blocks) x (row ranges): partial W per row
//   dense_wy_apply_kernel     range with DMMA, then X = -T^T (sum of the partials) and A2 += V X with DMMA
// The matrix lives in global memory; at the reference's sizes (5120 x 385 doubles = 15.8 MB) it is L2 resident.
// With a ColPiv right solver this is the first stage: ColPivHouseholderQR then runs on the M x M triangle (same P2, |R2|
// and x as on the tall matrix, the argument of the in-SM TSQR path) with dense_border.cuh's kernels.
#pragma once
#include <cooperative_groups.h>
#include "bd_wy.cuh"

namespace qrk {

struct DenseBlocked {
  double* A;          // N x ncols column-major: columns [0, M) the residual border, [M, ncols) right-hand sides
  long long ld, N;
  int M, ncols;
  double* tau;        // M
  double* T;          // 64 doubles per panel (column-major 8 x 8, upper triangular), panel p at T + 64 p
};

constexpr int kDbCluster = 8;      // CTAs per panel cluster (portable maximum)
constexpr int kDbThreads = 512;    // threads per CTA: 4096 rows per pass, RPT passes in registers

// Cluster-wide sum of nv <= 28 per-thread values (deterministic order: lanes, warps, CTA ranks — every CTA gets the
// bit-identical total).  Also fetches rank 0's pivot-row entries.  par: double-buffer parity (one cluster.sync per call).
constexpr size_t kDbPanelSmem = (size_t)28 * kDbThreads * sizeof(double);   // transposition buffer of the reductions

template <int NV>
__device__ __forceinline__ void cluster_reduce_vec(double (&v)[NV], int nv, int par, double* buf, double (*sall)[kDbCluster][28],
                                                   double (*sfin)[36], const double (*spiv)[8]) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster.block_rank();
  // CTA level through shared memory (value j of thread t at buf[j][t], then one warp sums a row): a shuffle tree over
  // nv values would cost 10 nv SHFL per warp, and the SHFL pipe issues one warp-instruction per clock per SM
#pragma unroll
  for (int j = 0; j < NV; j++) if (j < nv) buf[j * kDbThreads + tid] = v[j];
  __syncthreads();
  for (int j = warp; j < nv; j += kDbThreads / 32) {
    const double* row = buf + j * kDbThreads + lane;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int m = 0; m < kDbThreads / 32; m += 2) { s0 += row[32 * m]; s1 += row[32 * (m + 1)]; }
    double s = s0 + s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    // this CTA's total PUSHED into every CTA (lane l -> CTA l): the reads behind the barrier are local (fetching the 8
    // totals over DSMEM after the barrier cost 1.0k cycles per reduction in the trace of the sibling kernel)
    if (lane < kDbCluster) cluster.map_shared_rank(&sall[0][0][0], lane)[(par * kDbCluster + rank) * 28 + j] = s;
  }
  cluster.sync();                                  // every CTA's totals (and rank 0's pivot row) have landed here
  if (tid < nv) {
    double s = 0.0;
#pragma unroll
    for (int rk = 0; rk < kDbCluster; rk++) s += sall[par][rk][tid];          // fixed order: bit-identical on every CTA
    sfin[par][tid] = s;
  } else if (tid >= 28 && tid < 36) {
    sfin[par][tid] = spiv[par][tid - 28];
  }
  __syncthreads();
}

// Panel [k0, k0 + pw) of the matrix, rows [k0, N).  One cluster of 8 CTAs; thread t of CTA rank owns the rows
// k0 + rank * 512 + t + 4096 i, i < RPT.
#ifdef QRK_TRI_TRACE
__device__ long long g_panel_trace[16];
#define QRK_PANEL_CLK(i) do { if (k0 == 64 && threadIdx.x == 0 && blockIdx.x == 0) g_panel_trace[i] = clock64(); } while (0)
#else
#define QRK_PANEL_CLK(i) do { } while (0)
#endif

template <int RPT>
__global__ void __cluster_dims__(kDbCluster, 1, 1) __launch_bounds__(kDbThreads) dense_panel_kernel(DenseBlocked d, int k0, int pw) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(16) double sbuf[];      // kDbPanelSmem: 28 x 512 doubles
  __shared__ double sall[2][kDbCluster][28];      // the 8 CTAs' totals of the current reduction (pushed by their owners)
  __shared__ double sfin[2][36];
  __shared__ double spiv[2][8];                   // rank 0's pivot-row entries (pushed by rank 0)
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x;
  const int rank = (int)cluster.block_rank();
  const int row0 = rank * kDbThreads + tid;      // row (relative to k0) of pass 0
  QRK_PANEL_CLK(0);
  cluster.sync();                                 // every CTA has started: its shared memory may be written remotely
  QRK_PANEL_CLK(1);
  constexpr int PASS = kDbCluster * kDbThreads;
  double a[RPT][8];
#pragma unroll
  for (int i = 0; i < RPT; i++) {
    const long long r = (long long)k0 + row0 + (long long)PASS * i;
#pragma unroll
    for (int j = 0; j < 8; j++) a[i][j] = (r < d.N && j < pw) ? d.A[(long long)(k0 + j) * d.ld + r] : 0.0;
  }
  double tau_r[8];
#pragma unroll
  for (int c = 0; c < 8; c++) tau_r[c] = 0.0;
  QRK_PANEL_CLK(2);

#pragma unroll
  for (int c = 0; c < 8; c++) {
    if (c == 1) QRK_PANEL_CLK(3);
    if (c < pw) {                               // uniform
      const int par = c & 1;
      double part[8];
#pragma unroll
      for (int j = 0; j < 8; j++) part[j] = 0.0;
#pragma unroll
      for (int i = 0; i < RPT; i++) {
        const bool below = (row0 + PASS * i) > c;
        const double x = below ? a[i][c] : 0.0;
#pragma unroll
        for (int j = c; j < 8; j++) part[j - c] = fma(x, a[i][j], part[j - c]);
      }
      if (rank == 0 && tid < 32) {              // rank 0, thread c owns the pivot row: its entries into every CTA (lane l -> CTA l)
#pragma unroll
        for (int j = c; j < 8; j++) {
          const double pv = __shfl_sync(0xffffffffu, a[0][j], c);
          if (tid < kDbCluster) cluster.map_shared_rank(&spiv[0][0], tid)[par * 8 + j] = pv;
        }
      }
      cluster_reduce_vec<8>(part, 8 - c, par, sbuf, sall, sfin, spiv);
      const double tailSq = sfin[par][0], c0 = sfin[par][28 + c];
      double beta, tau, inv;                    // Eigen makeHouseholder (one short dependent chain: common.cuh)
      householder_scalars(c0, tailSq, false, beta, inv, tau);
      tau_r[c] = tau;
      double w[8];
#pragma unroll
      for (int j = c + 1; j < 8; j++) w[j] = tau * fma(inv, sfin[par][j - c], sfin[par][28 + j]);
#pragma unroll
      for (int i = 0; i < RPT; i++) {
        const int rel = row0 + PASS * i;
        if (rel > c) {
          const double vi = a[i][c] * inv;
          a[i][c] = vi;
#pragma unroll
          for (int j = c + 1; j < 8; j++) a[i][j] = fma(-w[j], vi, a[i][j]);
        } else if (rel == c) {
          a[i][c] = beta;
#pragma unroll
          for (int j = c + 1; j < 8; j++) a[i][j] -= w[j];
        }
      }
    }
  }

  QRK_PANEL_CLK(4);
  // ---- G = V^T V (strictly upper part), V unit lower trapezoidal: G[c][j] = V[j][c] + sum_{r > j} V[r][c] V[r][j]
  {
    double g[28];
#pragma unroll
    for (int e = 0; e < 28; e++) g[e] = 0.0;
#pragma unroll
    for (int i = 0; i < RPT; i++) {
      const int rel = row0 + PASS * i;
      int e = 0;
#pragma unroll
      for (int j = 1; j < 8; j++) {
        const double vj = (rel > j) ? a[i][j] : ((rel == j) ? 1.0 : 0.0);
#pragma unroll
        for (int c = 0; c < j; c++) { g[e] = fma(vj, a[i][c], g[e]); e++; }   // rel >= j > c: a[i][c] is V[rel][c]
      }
    }
    cluster_reduce_vec<28>(g, 28, pw & 1, sbuf, sall, sfin, spiv);
  }
  QRK_PANEL_CLK(5);
  if (rank == 0 && tid < 8) {                   // T[0:j, j] = -tau_j T[0:j, 0:j] G[0:j, j],  T[j][j] = tau_j
    // row i of T depends only on row i: lane i builds it in registers (one thread with an 8 x 8 local array and rolled
    // loops spent ~3 us per panel in local-memory round trips)
    const double* G = sfin[pw & 1];
    const int i = tid;
    double Trow[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      double val = 0.0;
      if (j < pw) {
        if (j == i) val = tau_r[j];
        else if (j > i) {
          double sum = 0.0;
#pragma unroll
          for (int l = 0; l < j; l++) if (l >= i) sum = fma(Trow[l], G[j * (j - 1) / 2 + l], sum);     // G[c][j] sits at G[e0 + c]
          val = -tau_r[j] * sum;
        }
      }
      Trow[j] = val;
    }
    double* Tg = d.T + 64 * (k0 / 8);
#pragma unroll
    for (int j = 0; j < 8; j++) Tg[i + 8 * j] = Trow[j];
    if (i < pw) {
      double t = 0.0;
#pragma unroll
      for (int j = 0; j < 8; j++) if (j == i) t = tau_r[j];
      d.tau[k0 + i] = t;
    }
  }
  QRK_PANEL_CLK(6);
#pragma unroll
  for (int i = 0; i < RPT; i++) {
    const long long r = (long long)k0 + row0 + (long long)PASS * i;
    if (r < d.N) {
#pragma unroll
      for (int j = 0; j < 8; j++) if (j < pw) d.A[(long long)(k0 + j) * d.ld + r] = a[i][j];
    }
  }
  QRK_PANEL_CLK(7);
  cluster.sync();                               // no CTA exits while its shared memory may still be read remotely
  QRK_PANEL_CLK(8);
}

// entry (r, j) of the unit lower trapezoidal V of panel k0 (stored below the diagonal of the panel columns)
__device__ __forceinline__ double wy_v(const DenseBlocked& d, int k0, int pw, long long r, int j) {
  if (r >= d.N || j >= pw) return 0.0;
  const long long rel = r - k0;
  if (rel < j) return 0.0;
  if (rel == j) return 1.0;
  return d.A[(long long)(k0 + j) * d.ld + r];
}

// Trailing update  A2 <- (I - V T^T V^T) A2  of the columns right of panel k0, rows [k0, N), as two launches so that BOTH
// the column blocks and the rows are spread over the GPU (a CTA's time is its number of dependent L2 round trips: with the
// rows split RS ways every CTA makes only a few):
//   dense_wy_w_kernel      grid (column blocks, RS): partial W = V^T A2 over the CTA's row range, DMMA, 4 rows per step
//   dense_wy_apply_kernel  grid (column blocks, RS): sums the RS partials in a fixed order, X = -T^T W, A2 += V X on its
//                          row range, DMMA, 8 rows per step
// first_block: column block offset (the look-ahead updates block 0 on its own stream).
constexpr int kWyWarps = 8;

__device__ __forceinline__ void wy_row_range(const DenseBlocked& d, int k0, int rs, int nrs, long long& lo, long long& hi) {
  const long long rows = d.N - k0;
  const long long chunk = ((rows + nrs - 1) / nrs + 7) & ~7LL;       // multiple of 8 rows
  lo = k0 + chunk * rs;
  hi = lo + chunk < d.N ? lo + chunk : d.N;
}

__global__ void __launch_bounds__(32 * kWyWarps) dense_wy_w_kernel(DenseBlocked d, int k0, int pw, int first_block, double* __restrict__ wpart) {
  __shared__ double sW[kWyWarps][64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = lane & 3, g = lane >> 2;
  const int blk = first_block + (int)blockIdx.x;
  const int col0 = k0 + pw + 8 * blk;
  const int ncb = (d.ncols - col0 < 8) ? (d.ncols - col0) : 8;
  long long lo, hi;
  wy_row_range(d, k0, (int)blockIdx.y, (int)gridDim.y, lo, hi);
  // D[m = reflector][n = column] += sum_k V[r0 + k][m] A2[r0 + k][n]
  double w0 = 0.0, w1 = 0.0;
  const double* colp = d.A + (long long)(col0 + (g < ncb ? g : 0)) * d.ld;
  for (long long r0 = lo + 4 * warp; r0 < hi; r0 += 32 * kWyWarps) {
    double va[8], vb[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {                // 16 independent loads in flight per lane
      const long long r = r0 + 4 * kWyWarps * u + q;
      const bool in = r < hi;
      va[u] = in ? wy_v(d, k0, pw, r, g) : 0.0;
      vb[u] = (in && g < ncb) ? colp[r] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) dmma884(w0, w1, va[u], vb[u]);
  }
  sW[warp][g + 8 * (2 * q)] = w0;               // W[m + 8 n]
  sW[warp][g + 8 * (2 * q + 1)] = w1;
  __syncthreads();
  if (tid < 64) {
    double s = 0.0;
#pragma unroll
    for (int wv = 0; wv < kWyWarps; wv++) s += sW[wv][tid];
    wpart[((size_t)blk * gridDim.y + blockIdx.y) * 64 + tid] = s;
  }
}

__global__ void __launch_bounds__(32 * kWyWarps) dense_wy_apply_kernel(DenseBlocked d, int k0, int pw, int first_block, const double* __restrict__ wpart) {
  __shared__ double sW[64];
  __shared__ double sT[64];
  __shared__ double sX[64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = lane & 3, g = lane >> 2;
  const int blk = first_block + (int)blockIdx.x;
  const int col0 = k0 + pw + 8 * blk;
  const int ncb = (d.ncols - col0 < 8) ? (d.ncols - col0) : 8;
  long long lo, hi;
  wy_row_range(d, k0, (int)blockIdx.y, (int)gridDim.y, lo, hi);
  if (tid < 64) {
    sT[tid] = d.T[64 * (k0 / 8) + tid];
    double s = 0.0;
    for (int rs = 0; rs < (int)gridDim.y; rs++) s += wpart[((size_t)blk * gridDim.y + rs) * 64 + tid];
    sW[tid] = s;
  }
  __syncthreads();
  if (tid < 64) {                               // X = -T^T W:  X[k][n] = -sum_{m <= k} T[m][k] W[m][n]
    const int k = tid & 7, n = tid >> 3;
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < 8; m++) if (m <= k) s = fma(sT[m + 8 * k], sW[m + 8 * n], s);
    sX[k + 8 * n] = -s;
  }
  __syncthreads();
  // A2 += V X, 8 rows per step: C[m = row][n = column], A = V (8 x 4, two k-steps), B = X
  const double b0 = sX[q + 8 * g], b1 = sX[4 + q + 8 * g];
  const int n0 = 2 * q, n1 = 2 * q + 1;
  double* c0p = d.A + (long long)(col0 + (n0 < ncb ? n0 : 0)) * d.ld;
  double* c1p = d.A + (long long)(col0 + (n1 < ncb ? n1 : 0)) * d.ld;
  for (long long r0 = lo + 8 * warp; r0 < hi; r0 += 32 * kWyWarps) {
    double v0[4], v1[4], c0[4], c1[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long r = r0 + 8 * kWyWarps * u + g;
      const bool in = r < hi;
      v0[u] = in ? wy_v(d, k0, pw, r, q) : 0.0;
      v1[u] = in ? wy_v(d, k0, pw, r, 4 + q) : 0.0;
      c0[u] = (in && n0 < ncb) ? c0p[r] : 0.0;
      c1[u] = (in && n1 < ncb) ? c1p[r] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long r = r0 + 8 * kWyWarps * u + g;
      dmma884(c0[u], c1[u], v0[u], b0);
      dmma884(c0[u], c1[u], v1[u], b1);
      if (r < hi) {
        if (n0 < ncb) c0p[r] = c0[u];
        if (n1 < ncb) c1p[r] = c1[u];
      }
    }
  }
}

// upper triangle (and the right-hand side's first M entries) of the factored residual -> tri (M x (M + nrhs), ld = M)
__global__ void __launch_bounds__(256) dense_extract_tri_kernel(const double* __restrict__ A, long long ld, int M, int ncols,
                                                                double* __restrict__ tri) {
  const long long total = (long long)M * ncols;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int c = (int)(e / M), r = (int)(e - (long long)c * M);
    tri[e] = (c >= M || r <= c) ? A[(long long)c * ld + r] : 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// Second stage with a ColPiv right solver: Eigen's ColPivHouseholderQR (first maximum of the LAWN-176 downdated norms,
// recompute test, nonzero-pivot threshold — the rules of dense_border.cuh) on the M x (M + nrhs) triangle, in ONE launch.
// One thread-block cluster of 8 CTAs keeps the whole matrix in shared memory, columns dealt round-robin (column j lives in
// CTA j mod 8), so the work stays balanced as the factorisation advances.  Per column step: local pivot candidates ->
// cluster.sync -> global first maximum; the owner of column k swaps it with the pivot column (through distributed shared
// memory when that lives in another CTA), builds the reflector and publishes it -> cluster.sync -> every CTA copies v
// over DSMEM and updates its own columns (one warp per column) and their norms.  Two hardware cluster barriers per step
// instead of two kernel launches.
// d: A = triangle (ld = N = M), Nrule = rows of the tall residual; needs (ceil((M+nrhs)/8) + 2) * M doubles of shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kTriThreads = 512;
#ifdef QRK_TRI_TRACE
__device__ long long g_tri_trace[16];
#define QRK_TRI_CLK(i) do { if (k == 64 && tid == 0 && rank == 0) g_tri_trace[i] = clock64(); } while (0)
#else
#define QRK_TRI_CLK(i) do { } while (0)
#endif

__host__ __device__ inline size_t tri_colpiv_smem_bytes(int M, int nrhs) {
  const size_t cpc = (size_t)(M + nrhs + kDbCluster - 1) / kDbCluster;
  return (cpc * M + 2 * (size_t)M + 2 * cpc + 16) * sizeof(double) + (cpc + 8) * sizeof(int);
}

__global__ void __cluster_dims__(kDbCluster, 1, 1) __launch_bounds__(kTriThreads) dense_tri_colpiv_kernel(DenseBorder d) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) double dyn[];
  __shared__ double sred[kTriThreads / 32];
  __shared__ double candv[2][kDbCluster];      // pivot candidates of the 8 CTAs (value, index), double buffered by step parity
  __shared__ int candi[2][kDbCluster];
  const int M = d.M, NC = M + d.nrhs, LDR = M;
  const int CPC = (NC + kDbCluster - 1) / kDbCluster;
  double* cols = dyn;                        // CPC columns of M rows: local column l is global column rank + 8 l
  double* vloc = cols + (size_t)CPC * LDR;   // this step's reflector, local copy
  double* upd = vloc + 2 * M;                // (M doubles after vloc are unused padding)                     // m_colNormsUpdated of the local columns
  double* dir = upd + CPC;                   // m_colNormsDirect
  double* hdr = dir + CPC;                   // [0] tau [1] beta of the published reflector, [2..3] pivot candidate value (parity), [4] max norm
  int* perm = reinterpret_cast<int*>(hdr + 16);
  int* candj = perm + CPC;                   // [2] pivot candidate index (parity)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kTriThreads / 32;
  const int rank = (int)cluster.block_rank();

  // ---- load the local columns, their norms
  for (int l = 0; l < CPC; l++) {
    const int j = rank + kDbCluster * l;
    for (int i = tid; i < M; i += kTriThreads) cols[(size_t)l * LDR + i] = (j < NC) ? d.A[(size_t)j * M + i] : 0.0;
  }
  __syncthreads();
  double lmax = 0.0;
  for (int l = warp; l < CPC; l += NW) {
    const int j = rank + kDbCluster * l;
    double s = 0.0;
    for (int i = lane; i < M; i += 32) s = fma(cols[(size_t)l * LDR + i], cols[(size_t)l * LDR + i], s);
    s = warp_sum(s);
    const double nrm = sqrt(s);
    if (lane == 0) { upd[l] = nrm; dir[l] = nrm; perm[l] = j; }
    if (j < M) lmax = fmax(lmax, nrm);
  }
  if (lane == 0) sred[warp] = lmax;
  __syncthreads();
  if (tid == 0) {
    double m = 0.0;
    for (int w = 0; w < NW; w++) m = fmax(m, sred[w]);
    hdr[4] = m;
  }
  cluster.sync();
  double gmax = 0.0;
  for (int rk = 0; rk < kDbCluster; rk++) gmax = fmax(gmax, cluster.map_shared_rank(hdr, rk)[4]);
  const double me = gmax * DBL_EPSILON;
  const double helper = me * me / (double)d.Nrule;               // threshold_helper
  const int size = (int)(d.Nrule < M ? d.Nrule : M);
  const int steps = size < M ? size : M;
  int nonzero_pivots = size;
  double maxpivot = 0.0;

  for (int k = 0; k < steps; k++) {
    const int par = k & 1;
    QRK_TRI_CLK(0);
    // ---- (1) local pivot candidate: first maximum of upd over the local columns j >= k
    if (warp == 0) {
      double bv = -1.0;
      int bj = 0x7fffffff;
      for (int l = lane; l < CPC; l += 32) {
        const int j = rank + kDbCluster * l;
        if (j >= k && j < M) { const double u = upd[l]; if (u > bv) { bv = u; bj = j; } }   // j ascending: the first maximum stays
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
      }
      // pushed into every CTA of the cluster (8 remote stores by 8 lanes) instead of being fetched by every thread after the
      // barrier: the reads behind the barrier are local (r02 trace: the 8 dependent DSMEM fetches cost 1.6k cycles per column)
      if (lane < kDbCluster) {
        cluster.map_shared_rank(&candv[0][0], lane)[par * kDbCluster + rank] = bv;
        cluster.map_shared_rank(&candi[0][0], lane)[par * kDbCluster + rank] = bj;
      }
    }
    QRK_TRI_CLK(1);
    cluster.sync();
    QRK_TRI_CLK(2);
    // ---- (2) global first maximum over the 8 candidates (local copies)
    int big = 0x7fffffff;
    double bigv = -1.0;
#pragma unroll
    for (int rk = 0; rk < kDbCluster; rk++) {
      const double ov = candv[par][rk];
      const int oj = candi[par][rk];
      if (ov > bigv || (ov == bigv && oj < big)) { bigv = ov; big = oj; }
    }
    if (nonzero_pivots == size && bigv * bigv < helper * (double)(d.Nrule - k)) nonzero_pivots = k;
    QRK_TRI_CLK(3);
    // ---- (3) owner of column k: swap with the pivot column (fused with the tail norm), reflector, pushed to every CTA
    const int owner = k & (kDbCluster - 1);
    if (rank == owner) {
      const int lk = k / kDbCluster;
      double* ck = cols + (size_t)lk * LDR;
      double tailSq = 0.0;
      if (big != k) {
        const int ob = big & (kDbCluster - 1), lb = big / kDbCluster;
        double* cb = cluster.map_shared_rank(cols, ob) + (size_t)lb * LDR;
        for (int i = tid; i < M; i += kTriThreads) {
          const double t = ck[i], x = cb[i];
          ck[i] = x; cb[i] = t;
          if (i > k) tailSq = fma(x, x, tailSq);
        }
        if (tid == 0) {
          double* rupd = cluster.map_shared_rank(upd, ob);
          double* rdir = cluster.map_shared_rank(dir, ob);
          int* rperm = cluster.map_shared_rank(perm, ob);
          double t = upd[lk]; upd[lk] = rupd[lb]; rupd[lb] = t;
          t = dir[lk]; dir[lk] = rdir[lb]; rdir[lb] = t;
          const int p = perm[lk]; perm[lk] = rperm[lb]; rperm[lb] = p;
        }
      } else {
        for (int i = k + 1 + tid; i < M; i += kTriThreads) tailSq = fma(ck[i], ck[i], tailSq);
      }
      tailSq = warp_sum(tailSq);
      if (lane == 0) sred[warp] = tailSq;
      __syncthreads();
      tailSq = 0.0;
#pragma unroll
      for (int w = 0; w < NW; w++) tailSq += sred[w];
      const double c0 = ck[k];
      double beta, tau, inv;
      householder_scalars(c0, tailSq, false, beta, inv, tau);      // Eigen's makeHouseholder; one short dependent chain (common.cuh)
      __syncthreads();                                 // every thread has read ck[k] and sred
      for (int i = k + 1 + tid; i < M; i += kTriThreads) {
        const double v = ck[i] * inv;
        ck[i] = v;
#pragma unroll
        for (int rk = 0; rk < kDbCluster; rk++) cluster.map_shared_rank(vloc, rk)[i - k] = v;
      }
      if (tid < kDbCluster) {
        double* rv = cluster.map_shared_rank(vloc, tid);
        double* rh = cluster.map_shared_rank(hdr, tid);
        rv[0] = 1.0; rh[0] = tau; rh[1] = beta;
      }
      if (tid == 0) { ck[k] = beta; d.tau[k] = tau; }
    }
    QRK_TRI_CLK(4);
    cluster.sync();                                    // the reflector has landed in every CTA's vloc / hdr
    QRK_TRI_CLK(5);
    const double tau = hdr[0], beta = hdr[1];
    if (fabs(beta) > maxpivot) maxpivot = fabs(beta);
    // ---- (5) H_k on the local columns right of k, LAWN-176 downdate of their norms.  A warp takes FOUR local columns at a
    // time through the same loops (four independent dependency chains; 49 local columns of a 384-column border are one round
    // of the 16 warps instead of two), and the four norm downdates run in four lanes side by side instead of one after the
    // other in lane 0.  r02 trace at k = 64: this phase 8.4k of the column's 15.7k cycles before.
    for (int lb = warp; lb < CPC; lb += 4 * NW) {
      int lq[4], jq[4];
      bool act[4], any = false;
      int lfirst = -1;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        lq[q] = lb + q * NW; jq[q] = rank + kDbCluster * lq[q];
        act[q] = lq[q] < CPC && jq[q] > k && jq[q] < NC;
        if (act[q] && lfirst < 0) lfirst = lq[q];
        any = any || act[q];
      }
      if (!any) continue;
      double* cp[4];
#pragma unroll
      for (int q = 0; q < 4; q++) cp[q] = cols + (size_t)(act[q] ? lq[q] : lfirst) * LDR;     // an idle slot re-reads an active column (never stored)
      double dot[4] = {0.0, 0.0, 0.0, 0.0};
      double w[4], nsq[4] = {0.0, 0.0, 0.0, 0.0}, ak[4] = {0.0, 0.0, 0.0, 0.0};
      {
        for (int i = k + lane; i < M; i += 32) {
          const double v = vloc[i - k];
#pragma unroll
          for (int q = 0; q < 4; q++) dot[q] = fma(v, cp[q][i], dot[q]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int q = 0; q < 4; q++) dot[q] += __shfl_xor_sync(0xffffffffu, dot[q], o);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) w[q] = tau * dot[q];
        for (int i = k + lane; i < M; i += 32) {
          const double v = vloc[i - k];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const double x = fma(-v, w[q], cp[q][i]);
            if (act[q]) cp[q][i] = x;
            if (i > k) nsq[q] = fma(x, x, nsq[q]); else ak[q] = x;
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 4; q++) nsq[q] += __shfl_xor_sync(0xffffffffu, nsq[q], o);
      }
      // lane q finishes column q (row k is lane 0's first element: its entry travels by shuffle)
      double my_ak = 0.0, my_nsq = 0.0;
      int my_l = 0, my_j = NC;
      bool my_act = false;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const double a0 = __shfl_sync(0xffffffffu, ak[q], 0);
        if (lane == q) { my_ak = a0; my_nsq = nsq[q]; my_l = lq[q]; my_j = jq[q]; my_act = act[q]; }
      }
      if (lane < 4 && my_act && my_j < M) {                      // (a right-hand side has no norm)
        const double u = upd[my_l];
        if (u != 0.0) {
          // Eigen's rule (ColPivHouseholderQR.h, LAWN 176) with correctly rounded fast reciprocal / square root (common.cuh:
          // tools/seed_accuracy: 0 ulp against IEEE on 16M samples); the values only feed pivot comparisons
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
    QRK_TRI_CLK(6);
    __syncthreads();
    QRK_TRI_CLK(7);
  }

  // ---- write back: packed factors, P2, rank bookkeeping
  for (int l = 0; l < CPC; l++) {
    const int j = rank + kDbCluster * l;
    if (j < NC) for (int i = tid; i < M; i += kTriThreads) d.A[(size_t)j * M + i] = cols[(size_t)l * LDR + i];
    if (j < M && tid == 0) d.perm[j] = perm[l];
  }
  if (rank == 0 && tid == 0) { d.scal[0] = helper; d.scal[1] = maxpivot; d.iscal[0] = nonzero_pivots; }
  cluster.sync();                                      // shared memory stays alive while peers may still read it
}

}  // namespace qrk
