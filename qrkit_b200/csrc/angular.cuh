// angular.cuh — block-angular QR  A = [J1 | J2]  with J1 block diagonal (small uniform blocks) and a
// dense border J2 of M2 <= 8 columns: the block-diagonal stage + border merge of
// BlockAngularSparseQR::factorize (reference src/QRKit/BlockAngularSparseQR.h:459-514):
//     m_leftSolver.compute(J1)                                   (:472)
//     J2' = Q1^T J2 ; Atop = J2'[0:m1], Abot = J2'[m1:n]          (solveRightBlock :361-369)
//     rightSolver.compute(Abot) -> R2, P2                         (:368)
//     R = [R1, Atop P2; 0, R2]                                    (makeR :285-308)
// and of _solve_impl (:203-227) / the Q^T v product (:607-625).
//
// B200 design (one pass over the blocks, nothing of size n x n or n x m2 is ever re-read):
//   K1  angular_factor_kernel   one thread per diagonal block: Householder QR of the block in
//       registers, its reflectors applied at once to the block's rows of [J2 | b]; the top c rows go
//       to Atop / y1, the r-c residual rows never leave the SM: they are folded into a per-thread
//       M2 x (M2+1) triangle (TSQR leaf), threads are merged by a warp-shuffle butterfly, warps
//       through shared memory: one triangle per CTA goes to HBM (a few KB in total).
//   K2  angular_root_kernel     TSQR root: merges the CTA triangles (and, multi-GPU, the per-GPU
//       triangles gathered over NCCL), then runs Eigen's ColPivHouseholderQR rule on the M2 x M2
//       triangle itself (column norms of R_t equal those of Abot, so P2 and |R2| are the reference's),
//       rank rule, x2 = R2^-1 z.
//   K3  angular_backsolve_kernel  x1_i = P1 R1_i^-1 (y1_i - Atop_i x2), one thread per block.
// The reference has no TSQR: it runs a dense ColPiv QR on the tall (n-m1) x m2 matrix (:368); R2 agrees
// up to row signs, P2 and x agree (SURVEY §7 hard part 3).
#pragma once
#include "bd_small.cuh"

namespace qrk {

// Packed upper-trapezoidal M2 x (M2+1) triangle [R_t | z], stored by rows: row k holds columns k..M2.
template <int M2>
struct Tri {
  static constexpr int W = M2 + 1;
  static constexpr int N = W * (W + 1) / 2 - 1;
  __host__ __device__ static constexpr int idx(int k, int j) { return k * W - k * (k - 1) / 2 + (j - k); }
};

// Fold P dense rows (each W = M2+1 entries) into the triangle: Householder on [T_kk; w_0k .. w_{P-1,k}].
template <int M2, int P>
__device__ __forceinline__ void fold_rows(double (&T)[Tri<M2>::N], double (&w)[P][M2 + 1]) {
  using TR = Tri<M2>;
#pragma unroll
  for (int k = 0; k < M2; k++) {
    double tailSq = 0.0;
#pragma unroll
    for (int p = 0; p < P; p++) tailSq = fma(w[p][k], w[p][k], tailSq);
    const double c0 = T[TR::idx(k, k)];
    const bool degenerate = tailSq <= DBL_MIN;
    double norm;
    const double rnorm = fast_rsqrt(fma(c0, c0, tailSq), norm);
    double beta = (c0 >= 0.0) ? -norm : norm;
    const double ib = (c0 >= 0.0) ? -rnorm : rnorm;
    double inv = fast_rcp(c0 - beta);
    double tau = (beta - c0) * ib;
    if (degenerate) { inv = 0.0; tau = 0.0; beta = c0; }
    T[TR::idx(k, k)] = beta;
    double v[P];
#pragma unroll
    for (int p = 0; p < P; p++) v[p] = w[p][k] * inv;
#pragma unroll
    for (int j = k + 1; j <= M2; j++) {
      double dot = T[TR::idx(k, j)];
#pragma unroll
      for (int p = 0; p < P; p++) dot = fma(v[p], w[p][j], dot);
      dot *= tau;
      T[TR::idx(k, j)] -= dot;
#pragma unroll
      for (int p = 0; p < P; p++) w[p][j] = fma(-v[p], dot, w[p][j]);
    }
  }
}

// Fold another triangle S (same packing) into T, exploiting that row p of S is zero left of column p.
template <int M2>
__device__ __forceinline__ void fold_tri(double (&T)[Tri<M2>::N], double (&S)[Tri<M2>::N]) {
  using TR = Tri<M2>;
#pragma unroll
  for (int k = 0; k < M2; k++) {
    double tailSq = 0.0;
#pragma unroll
    for (int p = 0; p <= k; p++) tailSq = fma(S[TR::idx(p, k)], S[TR::idx(p, k)], tailSq);
    const double c0 = T[TR::idx(k, k)];
    const bool degenerate = tailSq <= DBL_MIN;
    double norm;
    const double rnorm = fast_rsqrt(fma(c0, c0, tailSq), norm);
    double beta = (c0 >= 0.0) ? -norm : norm;
    const double ib = (c0 >= 0.0) ? -rnorm : rnorm;
    double inv = fast_rcp(c0 - beta);
    double tau = (beta - c0) * ib;
    if (degenerate) { inv = 0.0; tau = 0.0; beta = c0; }
    T[TR::idx(k, k)] = beta;
    double v[M2];
#pragma unroll
    for (int p = 0; p <= k; p++) v[p] = S[TR::idx(p, k)] * inv;
#pragma unroll
    for (int j = k + 1; j <= M2; j++) {
      double dot = T[TR::idx(k, j)];
#pragma unroll
      for (int p = 0; p <= k; p++) dot = fma(v[p], S[TR::idx(p, j)], dot);
      dot *= tau;
      T[TR::idx(k, j)] -= dot;
#pragma unroll
      for (int p = 0; p <= k; p++) S[TR::idx(p, j)] = fma(-v[p], dot, S[TR::idx(p, j)]);
    }
  }
}

// Warp butterfly: after the call lane 0 holds the QR-merge of all 32 lanes' triangles.
template <int M2>
__device__ __forceinline__ void warp_merge_tri(double (&T)[Tri<M2>::N]) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double S[Tri<M2>::N];
#pragma unroll
    for (int i = 0; i < Tri<M2>::N; i++) S[i] = __shfl_xor_sync(0xffffffffu, T[i], o);
    fold_tri<M2>(T, S);
  }
}

// CTA merge: result valid in thread 0.  scratch: NWARPS * Tri::N doubles of shared memory.
template <int M2, int NWARPS>
__device__ __forceinline__ void cta_merge_tri(double (&T)[Tri<M2>::N], double* scratch) {
  constexpr int N = Tri<M2>::N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  warp_merge_tri<M2>(T);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; i++) scratch[warp * N + i] = T[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < N; i++) T[i] = (lane < NWARPS) ? scratch[lane * N + i] : 0.0;
#pragma unroll
    for (int o = 1; o < NWARPS; o <<= 1) {
      double S[N];
#pragma unroll
      for (int i = 0; i < N; i++) S[i] = __shfl_xor_sync(0xffffffffu, T[i], o);
      fold_tri<M2>(T, S);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K1: factorize the diagonal blocks, apply Q_i^T to [J2 | b], emit Atop / y1, reduce the residual rows.
//   J2   : n x M2 column-major, leading dimension ldj (rows of block i: i*R .. i*R+R-1)
//   b    : n (may be null: a zero rhs rides along)
//   atop : m1 x M2 column-major (leading dimension m1);  y1: m1
//   abot : optional (n-m1) x (M2+1) column-major panel [Abot | Q1^T b bottom] (leading dimension n-m1),
//          kept for later solve(b') calls; null in the fused compute+solve path
//   partials : gridDim.x triangles (Tri<M2>::N doubles each)
// ---------------------------------------------------------------------------------------------
template <int R, int C, int TPB>
struct AngularSmem {
  static constexpr int SA = Group<R * C>::stride;
  static constexpr int ST = Group<C>::stride;
  static constexpr int SP = GroupI32<C>::stride;
  static constexpr int offA = 0;
  static constexpr int offT = offA + TPB * SA;
  static constexpr int offP = offT + TPB * ST;
  template <int M2>
  static constexpr size_t bytes() {
    size_t perm_bytes = ((size_t)TPB * SP * 4 + 15) & ~(size_t)15;
    return (size_t)offP * 8 + perm_bytes + (size_t)(TPB / 32) * Tri<M2>::N * 8;
  }
};

template <int R, int C, bool PIV, int M2, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
angular_factor_kernel(const double* A_in, double* packed, double* __restrict__ tau_out, int* __restrict__ perm_out,
                      const double* __restrict__ J2, long long ldj, const double* __restrict__ b,
                      double* __restrict__ atop, double* __restrict__ y1, double* __restrict__ abot,
                      double* __restrict__ partials, long long nb) {
  static_assert(R > C, "the border merge needs residual rows (r > c)");
  using L = AngularSmem<R, C, TPB>;
  constexpr int M1 = R - C, W = M2 + 1;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem + L::offA;
  double* sT = smem + L::offT;
  int* sP = reinterpret_cast<int*>(smem + L::offP);
  double* sTri = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(sP) + (((size_t)TPB * L::SP * 4 + 15) & ~(size_t)15));
  const int t = threadIdx.x;
  const long long m1 = nb * C, nres = nb * M1;

  double T[Tri<M2>::N];
#pragma unroll
  for (int i = 0; i < Tri<M2>::N; i++) T[i] = 0.0;

  const long long ntiles = (nb + TPB - 1) / TPB;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long tile0 = tile * TPB;
    const int count = (int)((nb - tile0 < TPB) ? (nb - tile0) : TPB);
    __syncthreads();                       // previous tile's stage_out has drained the staging buffers
    stage_in_async<R * C, L::SA, TPB>(sA, A_in + tile0 * (R * C), count);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (t < count) {
      const long long blk = tile0 + t;
      double a[R * C], tau[C], inv_diag[C], dummy[R];
      int perm[C];
      load_group<R * C>(a, sA + t * L::SA);
      BlockQR<R, C, PIV, false>::run(a, tau, inv_diag, perm, dummy);
      store_group<R * C>(sA + t * L::SA, a);
      store_group<C>(sT + t * L::ST, tau);
      if (PIV) {
#pragma unroll
        for (int j = 0; j < C; j++) sP[t * L::SP + j] = (int)(blk * C) + perm[j];
      }
      // border rows of this block: column by column through the reflectors
      double w[M1][W];
#pragma unroll
      for (int j = 0; j < W; j++) {
        double col[R];
        if (j < M2) {
          const double* src = J2 + (long long)j * ldj + blk * R;
#pragma unroll
          for (int i = 0; i < R; i++) col[i] = __ldg(src + i);
        } else {
#pragma unroll
          for (int i = 0; i < R; i++) col[i] = b ? __ldg(b + blk * R + i) : 0.0;
        }
        apply_qt_chain<R, C>(a, tau, col);
        double* top = (j < M2) ? atop + (long long)j * m1 + blk * C : y1 + blk * C;
#pragma unroll
        for (int i = 0; i < C; i++) top[i] = col[i];
#pragma unroll
        for (int i = 0; i < M1; i++) w[i][j] = col[C + i];
        if (abot) {
          double* dst = abot + (long long)j * nres + blk * M1;
#pragma unroll
          for (int i = 0; i < M1; i++) dst[i] = col[C + i];
        }
      }
      fold_rows<M2, M1>(T, w);
    }
    __syncthreads();
    stage_out<R * C, L::SA, TPB>(packed + tile0 * (R * C), sA, count);
    stage_out<C, L::ST, TPB>(tau_out + tile0 * C, sT, count);
    if (PIV) stage_out_i32<C, L::SP, TPB>(perm_out + tile0 * C, sP, count);
  }
  __syncthreads();
  cta_merge_tri<M2, TPB / 32>(T, sTri);
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < Tri<M2>::N; i++) partials[(long long)blockIdx.x * Tri<M2>::N + i] = T[i];
  }
}

// Solve-only variant of K1 on a stored factorisation: y1 = (Q1^T b) top, residual rows of b appended to
// the stored Abot panel (column M2) and the TSQR redone over [Abot | b_bot] (deterministic: same R_t).
template <int R, int C, int M2, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
angular_rhs_kernel(const double* __restrict__ packed, const double* __restrict__ tau_in, const double* __restrict__ b,
                   double* __restrict__ y1, double* __restrict__ abot, double* __restrict__ partials, long long nb) {
  constexpr int M1 = R - C, W = M2 + 1;
  constexpr int SA = Group<R * C>::stride, ST = Group<C>::stride;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sT = smem + TPB * SA;
  double* sTri = sT + TPB * ST;
  const int t = threadIdx.x;
  const long long nres = nb * M1;
  double T[Tri<M2>::N];
#pragma unroll
  for (int i = 0; i < Tri<M2>::N; i++) T[i] = 0.0;
  const long long ntiles = (nb + TPB - 1) / TPB;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long tile0 = tile * TPB;
    const int count = (int)((nb - tile0 < TPB) ? (nb - tile0) : TPB);
    __syncthreads();
    stage_in_async<R * C, SA, TPB>(sA, packed + tile0 * (R * C), count);
    stage_in_async<C, ST, TPB>(sT, tau_in + tile0 * C, count);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (t < count) {
      const long long blk = tile0 + t;
      double a[R * C], tau[C], col[R];
      load_group<R * C>(a, sA + t * SA);
      load_group<C>(tau, sT + t * ST);
#pragma unroll
      for (int i = 0; i < R; i++) col[i] = __ldg(b + blk * R + i);
      apply_qt_chain<R, C>(a, tau, col);
#pragma unroll
      for (int i = 0; i < C; i++) y1[blk * C + i] = col[i];
      double w[M1][W];
#pragma unroll
      for (int j = 0; j < M2; j++) {
#pragma unroll
        for (int i = 0; i < M1; i++) w[i][j] = __ldg(abot + (long long)j * nres + blk * M1 + i);
      }
#pragma unroll
      for (int i = 0; i < M1; i++) {
        w[i][M2] = col[C + i];
        abot[(long long)M2 * nres + blk * M1 + i] = col[C + i];
      }
      fold_rows<M2, M1>(T, w);
    }
  }
  __syncthreads();
  cta_merge_tri<M2, TPB / 32>(T, sTri);
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < Tri<M2>::N; i++) partials[(long long)blockIdx.x * Tri<M2>::N + i] = T[i];
  }
}

// ---------------------------------------------------------------------------------------------
// K2: TSQR root.  mode 0: merge `count` triangles into out_tri (the per-GPU triangle; multi-GPU step 1).
//     mode 1: merge, then ColPiv + rank + x2 (single GPU, or after the NCCL all-gather of G triangles).
//   root : [R2 (M2*M2 col-major) | z2 (M2) | y2 = R2^-1 z2 in pivoted order (M2) | x2 (M2, unpivoted)] doubles
//   root_i: [P2 (M2) | rank2]
// ---------------------------------------------------------------------------------------------
template <int M2, int TPB>
__global__ void __launch_bounds__(TPB)
angular_root_kernel(const double* __restrict__ tris, int count, int mode, double* __restrict__ out_tri,
                    double* __restrict__ root, int* __restrict__ root_i, int keep_rhs_only) {
  using TR = Tri<M2>;
  constexpr int N = TR::N;
  __shared__ double scratch[(TPB / 32) * N];
  double T[N];
#pragma unroll
  for (int i = 0; i < N; i++) T[i] = 0.0;
  for (int q = threadIdx.x; q < count; q += TPB) {
    double S[N];
#pragma unroll
    for (int i = 0; i < N; i++) S[i] = tris[(long long)q * N + i];
    fold_tri<M2>(T, S);
  }
  cta_merge_tri<M2, TPB / 32>(T, scratch);
  if (threadIdx.x != 0) return;
  if (mode == 0) {
#pragma unroll
    for (int i = 0; i < N; i++) out_tri[i] = T[i];
    return;
  }
  // ColPivHouseholderQR of the M2 x M2 triangle R_t with z as the right-hand side
  double a[M2 * M2], tau[M2], inv_diag[M2], z[M2];
  int perm[M2];
#pragma unroll
  for (int j = 0; j < M2; j++) {
#pragma unroll
    for (int i = 0; i < M2; i++) a[j * M2 + i] = (i <= j) ? T[TR::idx(i, j)] : 0.0;
    z[j] = T[TR::idx(j, M2)];
  }
  BlockQR<M2, M2, true, true>::run(a, tau, inv_diag, perm, z);
  // rank as Eigen's ColPivHouseholderQR::rank(): |R_ii| > |maxpivot| * eps * diagonalSize
  double maxpivot = 0.0;
#pragma unroll
  for (int i = 0; i < M2; i++) maxpivot = fmax(maxpivot, fabs(a[i * M2 + i]));
  const double thresh = maxpivot * (DBL_EPSILON * (double)M2);
  int rank = 0;
#pragma unroll
  for (int i = 0; i < M2; i++) rank += (fabs(a[i * M2 + i]) > thresh) ? 1 : 0;
  if (!keep_rhs_only) {
#pragma unroll
    for (int j = 0; j < M2; j++) {
#pragma unroll
      for (int i = 0; i < M2; i++) root[j * M2 + i] = (i <= j) ? a[j * M2 + i] : 0.0;
      root_i[j] = perm[j];
    }
    root_i[M2] = rank;
  }
  double y[M2];
#pragma unroll
  for (int j = 0; j < M2; j++) { root[M2 * M2 + j] = z[j]; y[j] = z[j]; }
  // y[0:rank] = R2[0:rank,0:rank]^-1 z[0:rank]; the rest zero (_solve_impl :216-217)
#pragma unroll
  for (int j = M2 - 1; j >= 0; --j) {
    double s = y[j];
#pragma unroll
    for (int k = j + 1; k < M2; k++) s = fma(-a[k * M2 + j], y[k], s);
    y[j] = (j < rank) ? s * inv_diag[j] : 0.0;
  }
#pragma unroll
  for (int j = 0; j < M2; j++) root[M2 * M2 + M2 + j] = y[j];
#pragma unroll
  for (int j = 0; j < M2; j++) {
    root[M2 * M2 + 2 * M2 + perm[j]] = y[j];   // x2[P2[j]] = y_j (dest = P_c * y, :222)
  }
}

// ---------------------------------------------------------------------------------------------
// K3: x1 = P1 R1^-1 (y1 - Atop x2); also copies x2 to x[m1 + .].
// ---------------------------------------------------------------------------------------------
template <int R, int C, int M2, bool PERM, int TPB>
__global__ void __launch_bounds__(TPB)
angular_backsolve_kernel(const double* __restrict__ packed, const int* __restrict__ perm, const double* __restrict__ atop,
                         const double* __restrict__ y1, const double* __restrict__ root, double* __restrict__ x,
                         long long nb) {
  constexpr int SA = Group<R * C>::stride;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  const int t = threadIdx.x;
  const long long m1 = nb * C;
  double x2[M2];
#pragma unroll
  for (int j = 0; j < M2; j++) x2[j] = root[M2 * M2 + 2 * M2 + j];
  const long long tile0 = (long long)blockIdx.x * TPB;
  const int count = (int)((nb - tile0 < TPB) ? (nb - tile0) : TPB);
  if (blockIdx.x == 0 && t < M2) x[m1 + t] = root[M2 * M2 + 2 * M2 + t];
  stage_in_async<R * C, SA, TPB>(sA, packed + tile0 * (R * C), count);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (t >= count) return;
  const long long blk = tile0 + t;
  double a[R * C], y[R];
  load_group<R * C>(a, sA + t * SA);
#pragma unroll
  for (int k = 0; k < C; k++) y[k] = __ldg(y1 + blk * C + k);
#pragma unroll
  for (int j = 0; j < M2; j++) {
#pragma unroll
    for (int k = 0; k < C; k++) y[k] = fma(-__ldg(atop + (long long)j * m1 + blk * C + k), x2[j], y[k]);
  }
  double inv_diag[C];
#pragma unroll
  for (int k = 0; k < C; k++) inv_diag[k] = fast_rcp(a[k * R + k]);
  back_substitute<R, C>(a, inv_diag, y);
#pragma unroll
  for (int k = 0; k < C; k++) {
    const long long dst = PERM ? (long long)perm[blk * C + k] : blk * C + k;
    x[dst] = y[k];
  }
}

}  // namespace qrk
