// bd_small.cuh — batched FP64 Householder QR of small uniform diagonal blocks, ONE THREAD PER BLOCK.
//
// Replaces the serial per-block loop of BlockDiagonalSparseQR::factorize
// (reference src/QRKit/BlockDiagonalSparseQR.h:432-526: BlockQRSolver::compute per block, explicit
// r x r Q_i, per-coefficient insertBack) and, fused, the sparse Q^T b product + sparse triangular
// solve of _solve_impl (:258-280).  The arithmetic restated per block is Eigen's unblocked
// HouseholderQR / ColPivHouseholderQR (third party, not in the reference tree; the rules are stated in DESIGN.md).
//
// Why thread-per-block (SURVEY §7 hard part 1): an r x c block with r*c <= ~48 doubles fits in the
// registers of one thread, the whole factorisation unrolls into ~300 FP64 instructions with no
// shuffles and no divergence between neighbouring blocks, and 32 blocks per warp-instruction keep
// the SM under its issue budget so that the kernel stays HBM bound.  The price is the transposition
// between "consecutive blocks are consecutive in HBM" and "one block per thread"; it is paid in
// shared memory: coalesced 16-byte cp.async (LDGSTS) into a padded, bank-conflict-free layout, then
// 128-bit LDS per thread; results go back the same way (STS.128 -> coalesced STG.128).
//
// HBM traffic per block (fused factor+solve): read A 8rc + b 8r, write packed 8rc + tau 8c + x 8c
// (+ perm 4c) = 16rc + 8r + 16c (+4c) bytes — the algorithmic minimum.
#pragma once
#include "common.cuh"

namespace qrk {

// Eigen makeHouseholder on column K of the register-resident block (rows K..R-1).
// Returns tau; a[K*R+K] <- beta; a[K*R+i] (i>K) <- essential part; inv_beta <- 1/beta.
template <int R, int C, int K>
__device__ __forceinline__ double make_reflector_k(double (&a)[R * C], double& inv_beta) {
  double tailSq = 0.0;
#pragma unroll
  for (int i = K + 1; i < R; i++) tailSq = fma(a[K * R + i], a[K * R + i], tailSq);
  const double c0 = a[K * R + K];
  const bool degenerate = (K + 1 >= R) || (tailSq <= DBL_MIN);   // Eigen: tailSqNorm <= (numeric_limits::min)()
  double norm;
  const double rnorm = fast_rsqrt(fma(c0, c0, tailSq), norm);    // 1/||x||, ||x||
  double beta = (c0 >= 0.0) ? -norm : norm;                      // beta = -sign(x0) ||x||
  double ib = (c0 >= 0.0) ? -rnorm : rnorm;                      // 1/beta
  double inv = fast_rcp(c0 - beta);
  double tau = (beta - c0) * ib;
  if (degenerate) { inv = 0.0; tau = 0.0; beta = c0; ib = fast_rcp(c0); }
#pragma unroll
  for (int i = K + 1; i < R; i++) a[K * R + i] *= inv;
  a[K * R + K] = beta;
  inv_beta = ib;
  return tau;
}

// Apply H_K = I - tau v v^T (v = [1; ess]) to one column held in registers (rows K..R-1 of `col`).
template <int R, int C, int K>
__device__ __forceinline__ void apply_reflector_k(const double (&a)[R * C], double tau, double* col /* R entries */) {
  double tmp = col[K];
#pragma unroll
  for (int i = K + 1; i < R; i++) tmp = fma(a[K * R + i], col[i], tmp);
  tmp *= tau;
  col[K] -= tmp;
#pragma unroll
  for (int i = K + 1; i < R; i++) col[i] = fma(-a[K * R + i], tmp, col[i]);
}

// v <- Q^T v = H_{C-1} ... H_0 v (H_0 first) and v <- Q v = H_0 ... H_{C-1} v (H_{C-1} first)
template <int R, int C, int K = 0>
__device__ __forceinline__ void apply_qt_chain(const double (&a)[R * C], const double (&tau)[C], double (&v)[R]) {
  apply_reflector_k<R, C, K>(a, tau[K], v);
  if constexpr (K + 1 < ((R < C) ? R : C)) apply_qt_chain<R, C, K + 1>(a, tau, v);
}
template <int R, int C, int K = ((R < C) ? R : C) - 1>
__device__ __forceinline__ void apply_q_chain(const double (&a)[R * C], const double (&tau)[C], double (&v)[R]) {
  apply_reflector_k<R, C, K>(a, tau[K], v);
  if constexpr (K > 0) apply_q_chain<R, C, K - 1>(a, tau, v);
}

// One block: unpivoted (PIV=false) or Eigen-ColPiv (PIV=true) Householder QR in registers.
// rhs (R entries) is transformed alongside when RHS is true: rhs <- Q^T rhs.
// inv_diag[k] = 1 / R(k,k) (a by-product of the reflector; saves the divisions of the back substitution).
template <int R, int C, bool PIV, bool RHS>
struct BlockQR {
  static constexpr int NV = (R < C) ? R : C;

  template <int K>
  static __device__ __forceinline__ void step(double (&a)[R * C], double (&tau)[C], double (&inv_diag)[C], int (&perm)[C],
                                             double (&rhs)[R], double (&upd)[C], double (&dir)[C]) {
    if (PIV) {
      // first maximum of the downdated (squared) norms; strict '>' keeps the lowest index on ties
      int big = K;
      double bigv = upd[K];
#pragma unroll
      for (int j = K + 1; j < C; j++)
        if (upd[j] > bigv) { bigv = upd[j]; big = j; }
#pragma unroll
      for (int j = K + 1; j < C; j++) {
        if (big == j) {
#pragma unroll
          for (int i = 0; i < R; i++) { const double t = a[K * R + i]; a[K * R + i] = a[j * R + i]; a[j * R + i] = t; }
          { const double t = upd[K]; upd[K] = upd[j]; upd[j] = t; }
          { const double t = dir[K]; dir[K] = dir[j]; dir[j] = t; }
          { const int t = perm[K]; perm[K] = perm[j]; perm[j] = t; }
        }
      }
    }
    tau[K] = make_reflector_k<R, C, K>(a, inv_diag[K]);
#pragma unroll
    for (int j = K + 1; j < C; j++) apply_reflector_k<R, C, K>(a, tau[K], &a[j * R]);
    if (RHS) apply_reflector_k<R, C, K>(a, tau[K], rhs);
    if (PIV && K + 1 < NV) {
      // LAWN-176 norm downdate (Eigen ColPivHouseholderQR::computeInPlace) carried in SQUARED form:
      //   Eigen:  t = |a_kj|/upd_j; temp = (1+t)(1-t); temp2 = temp (upd_j/dir_j)^2;
      //           temp2 <= sqrt(eps) ? upd_j = dir_j = ||a[k+1:,j]|| : upd_j *= sqrt(temp)
      //   here :  U_j = upd_j^2, rho_j = (upd_j/dir_j)^2: temp = 1 - a_kj^2/U_j; temp2 = temp rho_j;
      //           temp2 <= sqrt(eps) ? (U_j = ||a[k+1:,j]||^2, rho_j = 1) : (U_j *= temp, rho_j *= temp)
      // The same quantities in exact arithmetic, with no square root and one reciprocal per column;
      // they only feed comparisons (pivot choice), never the factors.
      const double thr = 1.4901161193847656e-08;  // sqrt(DBL_EPSILON)
#pragma unroll
      for (int j = K + 1; j < C; j++) {
        const double akj = a[j * R + K];
        double temp = fmax(fma(-(akj * akj), fast_rcp(upd[j]), 1.0), 0.0);   // U_j = 0 -> NaN -> 0 -> recompute (= 0)
        const double temp2 = temp * dir[j];
        if (temp2 <= thr) {
          double s = 0.0;
#pragma unroll
          for (int i = K + 1; i < R; i++) s = fma(a[j * R + i], a[j * R + i], s);
          upd[j] = s;
          dir[j] = 1.0;
        } else {
          upd[j] *= temp;
          dir[j] *= temp;
        }
      }
    }
    if constexpr (K + 1 < NV) step<K + 1>(a, tau, inv_diag, perm, rhs, upd, dir);
  }

  static __device__ __forceinline__ void run(double (&a)[R * C], double (&tau)[C], double (&inv_diag)[C], int (&perm)[C],
                                            double (&rhs)[R]) {
    double upd[C], dir[C];
#pragma unroll
    for (int j = 0; j < C; j++) { perm[j] = j; tau[j] = 0.0; inv_diag[j] = 0.0; }
    if (PIV) {
#pragma unroll
      for (int j = 0; j < C; j++) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < R; i++) s = fma(a[j * R + i], a[j * R + i], s);
        upd[j] = s;      // squared norm U_j
        dir[j] = 1.0;    // rho_j = (upd_j/dir_j)^2
      }
    }
    step<0>(a, tau, inv_diag, perm, rhs, upd, dir);
  }
};

// y[0:C] <- R^-1 y[0:C]  (upper triangle of the packed block)
template <int R, int C>
__device__ __forceinline__ void back_substitute(const double (&a)[R * C], const double (&inv_diag)[C], double (&y)[R]) {
#pragma unroll
  for (int j = C - 1; j >= 0; --j) {
    double s = y[j];
#pragma unroll
    for (int k = j + 1; k < C; k++) s = fma(-a[k * R + j], y[k], s);
    y[j] = s * inv_diag[j];
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel 1: factorize (+ optionally fused Q^T b, back substitution and column permutation).
//   A_in  : nb blocks, block-COO layout (may alias packed_out: in-place factorisation)
//   packed: R in the upper triangle, essential Householder parts below (Eigen/LAPACK packing)
//   tau   : C per block;  perm: int32 C per block, GLOBAL column indices base_col + p_i(j)
//           (BlockDiagonalSparseQR.h:519-521); written only when PIV
//   b, x  : R resp. C per block (SOLVE only); x(base_col + p_i(j)) = y_j  (:275)
// ---------------------------------------------------------------------------------------------
template <int R, int C, int TPB>
struct SmallSmem {
  static constexpr int SA = Group<R * C>::stride;
  static constexpr int SB = Group<R>::stride;       // rhs in
  static constexpr int ST = Group<C>::stride;       // tau out; x out (aliases the rhs region, own stride)
  static constexpr int SP = GroupI32<C>::stride;
  static constexpr int offA = 0;
  static constexpr int offB = offA + TPB * SA;
  static constexpr int offT = offB + TPB * (SB > ST ? SB : ST);
  static constexpr int offP = offT + TPB * ST;      // in doubles
  static constexpr size_t bytes = (size_t)offP * 8 + (size_t)TPB * SP * 4;
};

template <int R, int C, bool PIV, bool SOLVE, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
bd_small_factor_kernel(const double* A_in, double* packed, double* __restrict__ tau_out,
                       int* __restrict__ perm_out, const double* __restrict__ b, double* __restrict__ x, long long nb,
                       long long block0) {
  static_assert(R >= C, "portrait blocks only (BlockDiagonalSparseQR.h:509-516 rejects landscape blocks)");
  using L = SmallSmem<R, C, TPB>;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem + L::offA;
  double* sB = smem + L::offB;
  double* sT = smem + L::offT;
  int* sP = reinterpret_cast<int*>(smem + L::offP);

  const long long tile0 = (long long)blockIdx.x * TPB;
  const int count = (int)((nb - tile0 < TPB) ? (nb - tile0) : TPB);
  const int t = threadIdx.x;

  stage_in_async<R * C, L::SA, TPB>(sA, A_in + tile0 * (R * C), count);
  if (SOLVE) stage_in_async<R, L::SB, TPB>(sB, b + tile0 * R, count);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  double a[R * C], rhs[R], tau[C], inv_diag[C];
  int perm[C];
  if (t < count) {
    load_group<R * C>(a, sA + t * L::SA);
    if (SOLVE) load_group<R>(rhs, sB + t * L::SB);
  }
  if (SOLVE) __syncthreads();   // every rhs is in registers: the region is reused for x with a different stride
  if (t < count) {
    BlockQR<R, C, PIV, SOLVE>::run(a, tau, inv_diag, perm, rhs);
    store_group<R * C>(sA + t * L::SA, a);
    store_group<C>(sT + t * L::ST, tau);
    if (PIV) {
      const int base_col = (int)((block0 + tile0 + t) * C);   // block0: first block of this launch (chunked host pipeline)
#pragma unroll
      for (int j = 0; j < C; j++) sP[t * L::SP + j] = base_col + perm[j];
    }
    if (SOLVE) {
      back_substitute<R, C>(a, inv_diag, rhs);
      double* xs = sB + t * L::ST;
      if (PIV) {
#pragma unroll
        for (int j = 0; j < C; j++) xs[perm[j]] = rhs[j];
      } else {
        double xr[C];
#pragma unroll
        for (int j = 0; j < C; j++) xr[j] = rhs[j];
        store_group<C>(xs, xr);
      }
    }
  }
  __syncthreads();
  stage_out<R * C, L::SA, TPB>(packed + tile0 * (R * C), sA, count);
  stage_out<C, L::ST, TPB>(tau_out + tile0 * C, sT, count);
  if (PIV) stage_out_i32<C, L::SP, TPB>(perm_out + tile0 * C, sP, count);
  if (SOLVE) stage_out<C, L::ST, TPB>(x + tile0 * C, sB, count);
}

// ---------------------------------------------------------------------------------------------
// Kernel 2: operations on an existing factorisation, nrhs right-hand sides, block kept in registers.
//   OP_SOLVE    : X = P R^-1 (Q^T B)[0:C]                     (_solve_impl :258-280)
//   OP_APPLY_QT : Y = Q^T B in the reference's Q index layout (matrixQ().transpose()*B, :235,:266)
//   OP_APPLY_Q  : Y = Q B (input in the Q index layout)
// Q layout (qformat): FullQ  : thin part at base_col+k (k<C), complement at n_cols + m1off + k (:455-470)
//                     BlockDiagonalQ : base_row + k (:483-491)
// ---------------------------------------------------------------------------------------------
enum { OP_SOLVE = 0, OP_APPLY_QT = 1, OP_APPLY_Q = 2 };

template <int R, int C, int TPB>
struct SmallOpSmem {
  static constexpr int SA = Group<R * C>::stride;
  static constexpr int SB = Group<R>::stride;
  static constexpr int ST = Group<C>::stride;
  static constexpr int SC = Group<(R - C > 0 ? R - C : 1)>::stride;
  static constexpr int offA = 0;
  static constexpr int offB = offA + TPB * SA;
  static constexpr int offT = offB + TPB * (SB > ST ? SB : ST);
  static constexpr int offC = offT + TPB * ST;
  static constexpr size_t bytes = (size_t)(offC + TPB * SC) * 8;
};

template <int R, int C, int OP, bool PERM, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
bd_small_op_kernel(const double* __restrict__ packed, const double* __restrict__ tau_in, const int* __restrict__ perm,
                   const double* __restrict__ B, long long ldb, double* __restrict__ X, long long ldx, int nrhs,
                   long long nb, long long n_cols, int full_q) {
  using L = SmallOpSmem<R, C, TPB>;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem + L::offA;
  double* sB = smem + L::offB;
  double* sT = smem + L::offT;
  double* sC = smem + L::offC;
  const long long tile0 = (long long)blockIdx.x * TPB;
  const int count = (int)((nb - tile0 < TPB) ? (nb - tile0) : TPB);
  const int t = threadIdx.x;
  constexpr int M1 = R - C;

  stage_in_async<R * C, L::SA, TPB>(sA, packed + tile0 * (R * C), count);
  stage_in_async<C, L::ST, TPB>(sT, tau_in + tile0 * C, count);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  double a[R * C], tau[C];
  int p[C];
  if (t < count) {
    load_group<R * C>(a, sA + t * L::SA);
    load_group<C>(tau, sT + t * L::ST);
#pragma unroll
    for (int j = 0; j < C; j++) p[j] = PERM ? perm[(tile0 + t) * C + j] - (int)((tile0 + t) * C) : j;
  }
  // many right-hand sides over few blocks (Q1^T J2 of a wide border): gridDim.y splits the rhs range
  const int rhs_per = (nrhs + (int)gridDim.y - 1) / (int)gridDim.y;
  const int rhs_lo = (int)blockIdx.y * rhs_per, rhs_hi = (rhs_lo + rhs_per < nrhs) ? rhs_lo + rhs_per : nrhs;
  for (int rhs_i = rhs_lo; rhs_i < rhs_hi; rhs_i++) {
    const double* Bc = B + (long long)rhs_i * ldb;
    double* Xc = X + (long long)rhs_i * ldx;
    __syncthreads();
    if (OP == OP_APPLY_Q && full_q) {
      stage_in_async<C, L::ST, TPB>(sT, Bc + tile0 * C, count);
      if (M1 > 0) stage_in_async<(M1 > 0 ? M1 : 1), L::SC, TPB>(sC, Bc + n_cols + tile0 * M1, count);
    } else {
      stage_in_async<R, L::SB, TPB>(sB, Bc + tile0 * R, count);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    double v[R];
    if (t < count) {
      if (OP == OP_APPLY_Q && full_q) {
#pragma unroll
        for (int k = 0; k < C; k++) v[k] = sT[t * L::ST + k];
#pragma unroll
        for (int k = 0; k < M1; k++) v[C + k] = sC[t * L::SC + k];
      } else {
        load_group<R>(v, sB + t * L::SB);
      }
    }
    if (OP == OP_SOLVE) __syncthreads();   // the rhs region is reused for x with a different stride
    if (t < count) {
      if (OP == OP_APPLY_Q) {
        // Q v = H_0 ... H_{C-1} v : last reflector first
        apply_q_chain<R, C>(a, tau, v);
      } else {
        apply_qt_chain<R, C>(a, tau, v);
      }
      if (OP == OP_SOLVE) {
        double inv_diag[C];
#pragma unroll
        for (int j = 0; j < C; j++) inv_diag[j] = 1.0 / a[j * R + j];
        back_substitute<R, C>(a, inv_diag, v);
        double* xs = sB + t * L::ST;
#pragma unroll
        for (int j = 0; j < C; j++) xs[PERM ? p[j] : j] = v[j];
      } else if (OP == OP_APPLY_QT && full_q) {
#pragma unroll
        for (int k = 0; k < C; k++) sT[t * L::ST + k] = v[k];
#pragma unroll
        for (int k = 0; k < M1; k++) sC[t * L::SC + k] = v[C + k];
      } else {
        store_group<R>(sB + t * L::SB, v);
      }
    }
    __syncthreads();
    if (OP == OP_SOLVE) {
      stage_out<C, L::ST, TPB>(Xc + tile0 * C, sB, count);
    } else if (OP == OP_APPLY_QT && full_q) {
      stage_out<C, L::ST, TPB>(Xc + tile0 * C, sT, count);
      if (M1 > 0) stage_out<(M1 > 0 ? M1 : 1), L::SC, TPB>(Xc + n_cols + tile0 * M1, sC, count);
    } else {
      stage_out<R, L::SB, TPB>(Xc + tile0 * R, sB, count);
    }
  }
}

}  // namespace qrk
