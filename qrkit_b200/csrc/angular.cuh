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

// L2 residency between K1 and K3 (see common.cuh): -DQRK_ANG_NOHINT builds the kernels without the cache policies (A/B)
#ifdef QRK_ANG_NOHINT
constexpr bool kAngHint = false;
#else
constexpr bool kAngHint = true;
#endif

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

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Cooperative warp merge of the 32 lanes' triangles: ONE Householder QR of the stacked (32*M2) x (M2+1)
// matrix, rows spread over the lanes (lane l owns the M2 rows of its triangle), reflector norms and the
// v^T A dot products are warp-shuffle reductions.  Lane 0's row k is the pivot row of column k, so the
// result triangle ends up in lane 0; the other lanes' triangles are consumed.  Compared with a pairwise
// butterfly (5 dependent 2-triangle merges) this has one fifth of the dependent rsqrt/rcp chains.
template <int M2>
__device__ __forceinline__ void warp_merge_tri(double (&T)[Tri<M2>::N]) {
  using TR = Tri<M2>;
  const bool lead = (threadIdx.x & 31) == 0;
#pragma unroll
  for (int k = 0; k < M2; k++) {
    // ONE reduction phase per column: the squared tail norm and the RAW dot products t^T a_j (t = the unscaled column
    // entries of the lanes below the pivot) are summed together, before the reflector scalars exist:
    //   v = [1; inv t],  tau v^T a_j = tau (pivot_j + inv t^T a_j) = w_j,  a_j -= (w_j inv) t
    double red[M2 + 1];
#pragma unroll
    for (int j = k; j <= M2; j++) {
      double d = 0.0;
#pragma unroll
      for (int p = 0; p <= k; p++) d = fma(T[TR::idx(p, k)], T[TR::idx(p, j)], d);
      red[j] = lead ? 0.0 : d;                    // lane 0 contributes only the pivot row
    }
    double piv[M2 + 1];
#pragma unroll
    for (int j = k; j <= M2; j++) piv[j] = __shfl_sync(0xffffffffu, T[TR::idx(k, j)], 0);
#pragma unroll
    for (int j = k; j <= M2; j++) red[j] = warp_sum_f64(red[j]);
    double beta, inv, tau;
    householder_scalars(piv[k], red[k], false, beta, inv, tau);
#pragma unroll
    for (int j = k + 1; j <= M2; j++) {
      const double w = tau * fma(inv, red[j], piv[j]);
      if (lead) {
        T[TR::idx(k, j)] -= w;
      } else {
        const double z = w * inv;
#pragma unroll
        for (int p = 0; p <= k; p++) T[TR::idx(p, j)] = fma(-z, T[TR::idx(p, k)], T[TR::idx(p, j)]);
      }
    }
    if (lead) T[TR::idx(k, k)] = beta;
  }
}

// CTA merge: result valid in thread 0.  scratch: NWARPS * Tri::N doubles of shared memory.
template <int M2, int NWARPS>
__device__ __forceinline__ void cta_merge_tri(double (&T)[Tri<M2>::N], double* scratch) {
  constexpr int N = Tri<M2>::N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  warp_merge_tri<M2>(T);
  if (NWARPS == 1) return;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; i++) scratch[warp * N + i] = T[i];
  }
  __syncthreads();
  if (warp == 0) {
    static_assert(NWARPS <= 32, "one triangle per lane in the second stage");
#pragma unroll
    for (int i = 0; i < N; i++) T[i] = (lane < NWARPS) ? scratch[lane * N + i] : 0.0;
    warp_merge_tri<M2>(T);
  }
}

// The same cooperative merge with NT triangles per lane (lane l owns NT * M2 rows of the stacked matrix): ONE warp merges
// 32 * NT triangles with the shuffle traffic and the dependent chain of a 32-triangle merge -- the per-lane dot products
// grow, the 5 reductions and reflector-scalar chains per merge do not.  Measured on the root kernel (tools/root_trace):
// 16 warps merging at once take 5.0k cycles for what one warp does in 2.7k (they queue on the SM's shuffle port), a 4-triangle
// merge takes 3.0k.  Result in lane 0's T[0].  (NT = 1 is warp_merge_tri; see MergeFan for what is shipped.)
template <int M2, int NT>
__device__ __forceinline__ void warp_merge_tri_multi(double (&T)[NT][Tri<M2>::N]) {
  using TR = Tri<M2>;
  const bool lead = (threadIdx.x & 31) == 0;
#pragma unroll
  for (int k = 0; k < M2; k++) {
    double red[M2 + 1];
#pragma unroll
    for (int j = k; j <= M2; j++) {
      double d = 0.0;
#pragma unroll
      for (int t = 0; t < NT; t++) {
        double dt = 0.0;
#pragma unroll
        for (int p = 0; p <= k; p++) dt = fma(T[t][TR::idx(p, k)], T[t][TR::idx(p, j)], dt);
        d += (t == 0 && lead) ? 0.0 : dt;         // lane 0's first triangle holds the pivot rows only
      }
      red[j] = d;
    }
    double piv[M2 + 1];
#pragma unroll
    for (int j = k; j <= M2; j++) piv[j] = __shfl_sync(0xffffffffu, T[0][TR::idx(k, j)], 0);
#pragma unroll
    for (int j = k; j <= M2; j++) red[j] = warp_sum_f64(red[j]);
    double beta, inv, tau;
    householder_scalars(piv[k], red[k], false, beta, inv, tau);
#pragma unroll
    for (int j = k + 1; j <= M2; j++) {
      const double w = tau * fma(inv, red[j], piv[j]);
      const double z = w * inv;
#pragma unroll
      for (int t = 0; t < NT; t++) {
        if (t == 0 && lead) {
          T[0][TR::idx(k, j)] -= w;
        } else {
#pragma unroll
          for (int p = 0; p <= k; p++) T[t][TR::idx(p, j)] = fma(-z, T[t][TR::idx(p, k)], T[t][TR::idx(p, j)]);
        }
      }
    }
    if (lead) T[0][TR::idx(k, k)] = beta;
  }
}

// CTA merge by ONE warp (NWARPS triangles per lane): all (NWARPS * 32) thread triangles travel through shared memory
// (all: NWARPS * 32 * Tri::N doubles, entry i of thread t at i * threads + t), warp 0 merges them.  Result in thread 0.
template <int M2, int NWARPS>
__device__ __forceinline__ void cta_merge_tri_one_warp(double (&T)[Tri<M2>::N], double* all) {
  constexpr int N = Tri<M2>::N, TH = NWARPS * 32;
#pragma unroll
  for (int i = 0; i < N; i++) all[i * TH + threadIdx.x] = T[i];
  __syncthreads();
  if (threadIdx.x < 32) {
    double Tm[NWARPS][N];
#pragma unroll
    for (int t = 0; t < NWARPS; t++)
#pragma unroll
      for (int i = 0; i < N; i++) Tm[t][i] = all[i * TH + t * 32 + threadIdx.x];
    warp_merge_tri_multi<M2, NWARPS>(Tm);
#pragma unroll
    for (int i = 0; i < N; i++) T[i] = Tm[0][i];
  }
}

// Triangles per lane in the root's first merge level.  1 = one triangle per thread, 512 threads: the measured best IN THE STEP
// (config 3: root 8.8 us, step 49.5 us).  4 per lane on 128 threads is 7 % faster when the root is timed alone
// (tools/root_trace: 12.2k against 13.0k cycles) but 4.4 us slower inside the step (ncu without cache control: root 13.2 us,
// step 53.6 us; not explained -- sharing the SM with the dependent K3 grid was excluded by giving the root the whole SM's
// shared memory), so it stays an experiment switch: -DQRK_ANG_ROOT_NT=4 -DQRK_ANG_ROOT_TPB=128.
template <int M2>
#ifdef QRK_ANG_ROOT_NT
struct MergeFan { static constexpr int NT = QRK_ANG_ROOT_NT; };
#else
struct MergeFan { static constexpr int NT = 1; };
#endif

// ---------------------------------------------------------------------------------------------
// K1: factorize the diagonal blocks, apply Q_i^T to [J2 | b], emit Atop / y1, reduce the residual rows.
//   J2   : n x M2 column-major, leading dimension ldj (rows of block i: i*R .. i*R+R-1)
//   b    : n (may be null: a zero rhs rides along)
//   atop : m1 x M2 column-major (leading dimension m1);  y1: m1
//   abot : optional (n-m1) x (M2+1) column-major panel [Abot | Q1^T b bottom] (leading dimension n-m1),
//          kept for later solve(b') calls; null in the fused compute+solve path
//   partials : gridDim.x triangles (Tri<M2>::N doubles each)
// Persistent CTAs; a tile is U*TPB blocks (thread t owns blocks t, t+TPB, ... of the tile, so the U*(r-c)
// residual rows of one thread are folded with ONE reflector per border column).  The tile's slices of A,
// of every J2 column and of b are contiguous in HBM and are brought in by cp.async (LDGSTS) into a
// NSTAGE-deep ring of shared-memory buffers: the loads of tile i+1 are in flight while tile i is computed.
// ---------------------------------------------------------------------------------------------
template <int R, int C, int M2, int TPB, int U, int NSTAGE>
struct AngularSmem {
  static constexpr int TILE = U * TPB;
  static constexpr int W = M2 + 1;
  static constexpr int SA = Group<R * C>::stride;
  static constexpr int SR = Group<R>::stride;
  static constexpr int ST = Group<C>::stride;
  static constexpr int SP = GroupI32<C>::stride;
  static constexpr int stage_doubles = TILE * (SA + W * SR);        // A tile + W border/rhs column slices
  static constexpr int offT = NSTAGE * stage_doubles;               // tau out
  static constexpr int offTri = offT + TILE * ST;
  static constexpr int offP = offTri + (TPB / 32) * Tri<M2>::N + ((TPB / 32) * Tri<M2>::N % 2);
  static constexpr size_t bytes = (size_t)offP * 8 + (size_t)TILE * SP * 4;
};

template <int R, int C, bool PIV, int M2, int TPB, int U, int NSTAGE, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
angular_factor_kernel(const double* A_in, double* packed, double* __restrict__ tau_out, int* __restrict__ perm_out,
                      const double* __restrict__ J2, long long ldj, const double* __restrict__ b,
                      double* __restrict__ atop, double* __restrict__ y1, double* __restrict__ abot,
                      double* __restrict__ partials, long long nb, int pld) {
  static_assert(R > C, "the border merge needs residual rows (r > c)");
  using L = AngularSmem<R, C, M2, TPB, U, NSTAGE>;
  constexpr int M1 = R - C, W = M2 + 1, TILE = L::TILE;
  extern __shared__ __align__(16) double smem[];
  double* sT = smem + L::offT;
  double* sTri = smem + L::offTri;
  int* sP = reinterpret_cast<int*>(smem + L::offP);
  const int t = threadIdx.x;
  const long long m1 = nb * C, nres = nb * M1;

  double T[Tri<M2>::N];
#pragma unroll
  for (int i = 0; i < Tri<M2>::N; i++) T[i] = 0.0;

  const long long ntiles = (nb + TILE - 1) / TILE;
  const uint64_t pol_in = kAngHint ? l2_policy_evict_first() : 0, pol_out = kAngHint ? l2_policy_evict_last() : 0;
  auto issue = [&](long long tile, int stage) {
    const long long tile0 = tile * TILE;
    const int count = (int)((nb - tile0 < TILE) ? (nb - tile0) : TILE);
    double* buf = smem + (size_t)stage * L::stage_doubles;
    if (kAngHint) {
      stage_in_async_hint<R * C, L::SA, TPB, TILE>(buf, A_in + tile0 * (R * C), count, pol_in);
#pragma unroll
      for (int j = 0; j < M2; j++)
        stage_in_async_hint<R, L::SR, TPB, TILE>(buf + TILE * L::SA + j * TILE * L::SR, J2 + (long long)j * ldj + tile0 * R, count, pol_in);
      if (b) stage_in_async_hint<R, L::SR, TPB, TILE>(buf + TILE * L::SA + M2 * TILE * L::SR, b + tile0 * R, count, pol_in);
    } else {
      stage_in_async<R * C, L::SA, TPB, TILE>(buf, A_in + tile0 * (R * C), count);
#pragma unroll
      for (int j = 0; j < M2; j++)
        stage_in_async<R, L::SR, TPB, TILE>(buf + TILE * L::SA + j * TILE * L::SR, J2 + (long long)j * ldj + tile0 * R, count);
      if (b) stage_in_async<R, L::SR, TPB, TILE>(buf + TILE * L::SA + M2 * TILE * L::SR, b + tile0 * R, count);
    }
  };

  long long tile = blockIdx.x;
  if (NSTAGE > 1 && tile < ntiles) issue(tile, 0);
  if (NSTAGE > 1) cp_async_commit();
  int it = 0;
  for (; tile < ntiles; tile += gridDim.x, it++) {
    const int stage = (NSTAGE > 1) ? (it & 1) : 0;
    const long long tile0 = tile * TILE;
    const int count = (int)((nb - tile0 < TILE) ? (nb - tile0) : TILE);
    double* buf = smem + (size_t)stage * L::stage_doubles;
    if (NSTAGE > 1) {
      const long long next = tile + gridDim.x;
      if (next < ntiles) issue(next, stage ^ 1);     // the other buffer was drained before the barrier ending the previous iteration
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      issue(tile, 0);
      cp_async_commit();
      cp_async_wait<0>();
    }
    __syncthreads();
    double w[U * M1][W];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int slot = u * TPB + t;
      if (slot < count) {
        const long long blk = tile0 + slot;
        double a[R * C], tau[C], inv_diag[C], dummy[R];
        int perm[C];
        load_group<R * C>(a, buf + slot * L::SA);
        BlockQR<R, C, PIV, false>::run(a, tau, inv_diag, perm, dummy);
        store_group<R * C>(buf + slot * L::SA, a);
        store_group<C>(sT + slot * L::ST, tau);
        if (PIV) {
#pragma unroll
          for (int j = 0; j < C; j++) sP[slot * L::SP + j] = (int)(blk * C) + perm[j];
        }
#pragma unroll
        for (int j = 0; j < W; j++) {
          double col[R];
          if (j < M2 || b) load_group<R>(col, buf + TILE * L::SA + j * TILE * L::SR + slot * L::SR);
          else {
#pragma unroll
            for (int i = 0; i < R; i++) col[i] = 0.0;
          }
          apply_qt_chain<R, C>(a, tau, col);
          double* top = (j < M2) ? atop + (long long)j * m1 + blk * C : y1 + blk * C;
#pragma unroll
          for (int i = 0; i < C; i++) {
            if (kAngHint) st_global_hint(top + i, col[i], pol_out);
            else top[i] = col[i];
          }
#pragma unroll
          for (int i = 0; i < M1; i++) w[u * M1 + i][j] = col[C + i];
          if (abot) {
            double* dst = abot + (long long)j * nres + blk * M1;
#pragma unroll
            for (int i = 0; i < M1; i++) dst[i] = col[C + i];
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < M1; i++)
#pragma unroll
          for (int j = 0; j < W; j++) w[u * M1 + i][j] = 0.0;
      }
    }
    fold_rows<M2, U * M1>(T, w);
    __syncthreads();
    if (kAngHint) {
      stage_out_hint<R * C, L::SA, TPB, TILE>(packed + tile0 * (R * C), buf, count, pol_out);
      stage_out_hint<C, L::ST, TPB, TILE>(tau_out + tile0 * C, sT, count, pol_out);
    } else {
      stage_out<R * C, L::SA, TPB, TILE>(packed + tile0 * (R * C), buf, count);
      stage_out<C, L::ST, TPB, TILE>(tau_out + tile0 * C, sT, count);
    }
    if (PIV) {
      for (int d = t; d < count * C; d += TPB) perm_out[tile0 * C + d] = sP[(d / C) * L::SP + (d % C)];
    }
    __syncthreads();
  }
  if (NSTAGE > 1) cp_async_wait<0>();
  cta_merge_tri<M2, TPB / 32>(T, sTri);
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < Tri<M2>::N; i++) partials[(size_t)i * pld + blockIdx.x] = T[i];     // component-major: the root's loads coalesce
  }
}

// K1 for 2 x 1 blocks (the ellipse-fit shape of BASELINE configs 1 and 3): every per-point slice -- the block, its two rows
// of each border column and of b -- is ONE aligned 16-byte vector and consecutive points are consecutive in HBM, so a warp's
// plain vector loads are already perfectly coalesced (512 contiguous bytes each).  No shared-memory staging, no barriers in
// the loop: thread t of the grid takes points t, t + sweep, ...; U of them are folded with ONE reflector per border column,
// which is what pays (the 5 dependent reflector-scalar chains of a fold are the latency that bounds the kernel: ncu shows
// half of the issue slots empty on fixed-latency dependencies; U = 4 halves their number per point against the staged
// kernel's U = 2, and the 7 U independent vector loads of an iteration are all in flight together).
// Measured (profiles/r02_angular_2x1_kernels_ab.txt): config 3 53.7 -> 49.4 us together with the L2 policies and the direct
// K3.  Tried and dropped: next-iteration prefetch through a private cp.async ring (50.1 us at U = 3: the load wait it
// removes is not what bounds the step) and prefetch.global.L2 (53.9 us).
// Requires 16-byte aligned A / J2 / b, an even ldj and 32-bit point indices (checked by the dispatcher; otherwise the
// staged kernel runs).
// fold_rows with the fused reflector scalars of householder_scalars (the reciprocal's seed is taken from the 20-bit norm while
// the Newton steps of the rsqrt still run): a shorter dependent chain per border column than fold_rows' separate rsqrt + rcp
template <int M2, int P>
__device__ __forceinline__ void fold_rows_fused(double (&T)[Tri<M2>::N], double (&w)[P][M2 + 1]) {
  using TR = Tri<M2>;
#pragma unroll
  for (int k = 0; k < M2; k++) {
    double tailSq = 0.0;
#pragma unroll
    for (int p = 0; p < P; p++) tailSq = fma(w[p][k], w[p][k], tailSq);
    double beta, inv, tau;
    householder_scalars(T[TR::idx(k, k)], tailSq, false, beta, inv, tau);
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

// One iteration of the direct kernel: U points of one thread.  FULL: every one of them exists (no predication at all).
template <bool PIV, int M2, int U, bool ABOT, bool FULL>
__device__ __forceinline__ void angular_direct_step(double (&T)[Tri<M2>::N], unsigned p0, unsigned sweep, unsigned nb, unsigned ld2,
                                                    const double2* A2, double2* packed2, double* __restrict__ tau_out,
                                                    int* __restrict__ perm_out, const double2* __restrict__ J2v,
                                                    const double2* __restrict__ b2, double* __restrict__ atop,
                                                    double* __restrict__ y1, double* __restrict__ abot, uint64_t pol_in,
                                                    uint64_t pol_out) {
  constexpr int W = M2 + 1;
  const double2 zero2 = make_double2(0.0, 0.0);
  double2 av[U], cv[U][W];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const unsigned p = p0 + u * sweep;
    const bool live = FULL || p < nb;
    av[u] = !live ? zero2 : kAngHint ? ld_global_hint(A2 + p, pol_in) : A2[p];
#pragma unroll
    for (int j = 0; j < M2; j++) cv[u][j] = !live ? zero2 : kAngHint ? ld_global_nc_hint(J2v + (j * ld2 + p), pol_in) : __ldg(J2v + (j * ld2 + p));
    cv[u][M2] = !(live && b2) ? zero2 : kAngHint ? ld_global_nc_hint(b2 + p, pol_in) : __ldg(b2 + p);
  }
  double w[U][W];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const unsigned p = p0 + u * sweep;
    if (FULL || p < nb) {
      double a[2] = {av[u].x, av[u].y}, tau[1], inv_diag[1], dummy[2];
      int perm[1];
      BlockQR<2, 1, PIV, false>::run(a, tau, inv_diag, perm, dummy);
      if (kAngHint) {
        st_global_hint(packed2 + p, make_double2(a[0], a[1]), pol_out);
        st_global_hint(tau_out + p, tau[0], pol_out);
      } else {
        packed2[p] = make_double2(a[0], a[1]);
        tau_out[p] = tau[0];
      }
      if (PIV) perm_out[p] = (int)p;
#pragma unroll
      for (int j = 0; j < W; j++) {
        double col[2] = {cv[u][j].x, cv[u][j].y};
        apply_qt_chain<2, 1>(a, tau, col);
        double* top = (j < M2) ? atop + (j * nb + p) : y1 + p;
        if (kAngHint) st_global_hint(top, col[0], pol_out);
        else *top = col[0];
        w[u][j] = col[1];
        if (ABOT) abot[j * nb + p] = col[1];
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; j++) w[u][j] = 0.0;
    }
  }
#if defined(QRK_ANG_NOFOLD)
  // measurement only (wrong results): the memory-side floor of the kernel, with the fold's dependent chains removed
#pragma unroll
  for (int u = 0; u < U; u++)
#pragma unroll
    for (int j = 0; j < W; j++) T[j] += w[u][j];
#elif defined(QRK_ANG_FOLD_SPLIT)
  fold_rows<M2, U>(T, w);
#else
  fold_rows_fused<M2, U>(T, w);
#endif
}

// 32-bit point indices: the dispatcher guarantees (M2 + 1) * max(nb, ldj / 2) < 2^32.
template <bool PIV, int M2, int TPB, int U, int MINB, bool ABOT>
__global__ void __launch_bounds__(TPB, MINB)
angular_factor_direct_kernel(const double* A_in, double* packed, double* __restrict__ tau_out, int* __restrict__ perm_out,
                             const double* __restrict__ J2, long long ldj, const double* __restrict__ b,
                             double* __restrict__ atop, double* __restrict__ y1, double* __restrict__ abot,
                             double* __restrict__ partials, long long nb64, int pld) {
  // Epilogue: the CTA's 128 thread triangles are merged by ONE warp, four per lane (cta_merge_tri_one_warp), for borders of
  // up to 5 columns: 3 merging warps per SM instead of 12 followed by a second level (config 3: 49.5 -> 48.7 us;
  // -DQRK_ANG_K1_TWO_LEVEL restores the two-level merge for A/B).  Wider borders keep the two-level merge (registers).
#ifdef QRK_ANG_K1_TWO_LEVEL
  constexpr bool ONE_WARP = false;
#else
  constexpr bool ONE_WARP = Tri<M2>::N <= 20;
#endif
  __shared__ double sTri[ONE_WARP ? TPB * Tri<M2>::N : (TPB / 32) * Tri<M2>::N];
  double T[Tri<M2>::N];
#pragma unroll
  for (int i = 0; i < Tri<M2>::N; i++) T[i] = 0.0;
  const unsigned nb = (unsigned)nb64, ld2 = (unsigned)(ldj >> 1), sweep = gridDim.x * TPB;
  const uint64_t pol_in = kAngHint ? l2_policy_evict_first() : 0, pol_out = kAngHint ? l2_policy_evict_last() : 0;
  const double2* A2 = reinterpret_cast<const double2*>(A_in);
  double2* packed2 = reinterpret_cast<double2*>(packed);
  const double2* J2v = reinterpret_cast<const double2*>(J2);
  const double2* b2 = reinterpret_cast<const double2*>(b);
  unsigned p0 = blockIdx.x * TPB + threadIdx.x;
  for (; p0 + (U - 1) * sweep < nb; p0 += sweep * U)
    angular_direct_step<PIV, M2, U, ABOT, true>(T, p0, sweep, nb, ld2, A2, packed2, tau_out, perm_out, J2v, b2, atop, y1, abot, pol_in, pol_out);
  if (p0 < nb)
    angular_direct_step<PIV, M2, U, ABOT, false>(T, p0, sweep, nb, ld2, A2, packed2, tau_out, perm_out, J2v, b2, atop, y1, abot, pol_in, pol_out);
  if constexpr (ONE_WARP) cta_merge_tri_one_warp<M2, TPB / 32>(T, sTri);
  else cta_merge_tri<M2, TPB / 32>(T, sTri);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < Tri<M2>::N; i++) partials[(size_t)i * pld + blockIdx.x] = T[i];
  }
}

// Solve-only variant of K1 on a stored factorisation: y1 = (Q1^T b) top, residual rows of b appended to
// the stored Abot panel (column M2) and the TSQR redone over [Abot | b_bot] (deterministic: same R_t).
template <int R, int C, int M2, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
angular_rhs_kernel(const double* __restrict__ packed, const double* __restrict__ tau_in, const double* __restrict__ b,
                   double* __restrict__ y1, double* __restrict__ abot, double* __restrict__ partials, long long nb, int pld) {
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
    for (int i = 0; i < Tri<M2>::N; i++) partials[(size_t)i * pld + blockIdx.x] = T[i];     // component-major: the root's loads coalesce
  }
}

// Solve-only K1 for 2 x 1 blocks (the reference's own calling pattern is compute(J) followed by solve(b)): the unstaged
// counterpart of angular_rhs_kernel -- per point one vector load each of the packed block and of b, the scalar tau and the
// M2 stored residual entries; U points folded with one reflector per border column; the CTA merged by one warp where the
// triangle allows (as in angular_factor_direct_kernel).  Same partial-triangle layout (component-major).
template <int M2, int TPB, int U, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
angular_rhs_direct_kernel(const double* __restrict__ packed, const double* __restrict__ tau_in, const double* __restrict__ b,
                          double* __restrict__ y1, double* __restrict__ abot, double* __restrict__ partials, long long nb64, int pld) {
  constexpr int W = M2 + 1;
  constexpr bool ONE_WARP = Tri<M2>::N <= 20;
  __shared__ double sTri[ONE_WARP ? TPB * Tri<M2>::N : (TPB / 32) * Tri<M2>::N];
  double T[Tri<M2>::N];
#pragma unroll
  for (int i = 0; i < Tri<M2>::N; i++) T[i] = 0.0;
  const unsigned nb = (unsigned)nb64, sweep = gridDim.x * TPB;
  const double2* P2 = reinterpret_cast<const double2*>(packed);
  const double2* B2 = reinterpret_cast<const double2*>(b);
  for (unsigned p0 = blockIdx.x * TPB + threadIdx.x; p0 < nb; p0 += sweep * U) {
    double2 av[U], bv[U];
    double tv[U], w[U][W];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const unsigned p = p0 + u * sweep;
      const bool live = p < nb;
      av[u] = live ? __ldg(P2 + p) : make_double2(0.0, 0.0);
      bv[u] = live ? __ldg(B2 + p) : make_double2(0.0, 0.0);
      tv[u] = live ? __ldg(tau_in + p) : 0.0;
#pragma unroll
      for (int j = 0; j < M2; j++) w[u][j] = live ? __ldg(abot + ((size_t)j * nb + p)) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const unsigned p = p0 + u * sweep;
      const double a[2] = {av[u].x, av[u].y}, tau[1] = {tv[u]};
      double col[2] = {bv[u].x, bv[u].y};
      apply_qt_chain<2, 1>(a, tau, col);
      w[u][M2] = (p < nb) ? col[1] : 0.0;
      if (p < nb) { y1[p] = col[0]; abot[(size_t)M2 * nb + p] = col[1]; }
    }
    fold_rows_fused<M2, U>(T, w);
  }
  if constexpr (ONE_WARP) cta_merge_tri_one_warp<M2, TPB / 32>(T, sTri);
  else cta_merge_tri<M2, TPB / 32>(T, sTri);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < Tri<M2>::N; i++) partials[(size_t)i * pld + blockIdx.x] = T[i];
  }
}

// ---------------------------------------------------------------------------------------------
// K2: TSQR root.  mode 0: merge `count` triangles into out_tri (the per-GPU triangle; multi-GPU step 1).
//     mode 1: merge, then ColPiv + rank + x2 (single GPU, or after the NCCL all-gather of G triangles).
//     mode 2: merge, exchange the per-GPU triangles with the peers over NVLink inside this kernel, merge, root.
//   root : [R2 (M2*M2 col-major) | z2 (M2) | y2 = R2^-1 z2 in pivoted order (M2) | x2 (M2, unpivoted)] doubles
//   root_i: [P2 (M2) | rank2]
// ---------------------------------------------------------------------------------------------
// Peer exchange of the per-GPU triangles inside the root kernel (mode 2): the "all-gather" of SURVEY 8e as plain stores
// into every peer's exchange buffer over NVLink (buffers mapped with cudaIpcOpenMemHandle), followed by a flag per
// (step parity, source rank).  Buffer layout (identical on every rank): [2][G][Tri::N] doubles, then [2][G] uint64 flags.
// A rank can run at most one step ahead of a peer (it waits for the peer's flag of the current step), so two parities
// suffice.  The spin is bounded (timeout_ns of %globaltimer, default 10 s, qrk_angular_p2p_set_timeout): on a timeout *err is
// set (sticky until the next attach) and the merged triangle is POISONED with NaN, so that x2, x1, R2 and the root record of
// this and every later step are NaN on the ranks that saw it — never a plausible-looking wrong answer; the host entry points
// that synchronise (qrk_synchronize, qrk_rank, host-memspace solves) report QRK_STATUS_PEER_TIMEOUT.
struct AngularXchg {
  double* const* peers = nullptr;   // device array of G pointers: peers[g] = rank g's exchange buffer as mapped here
  int world = 0, rank = 0;
  unsigned long long* seq = nullptr; // device-resident step counter (the kernel advances it: a captured launch replays correctly)
  int* err = nullptr;
  unsigned long long timeout_ns = 10000000000ull;
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int M2, int TPB>
__device__ __forceinline__ void merge_triangle_list(const double* tris, int count, double (&T)[Tri<M2>::N], double* scratch,
                                                    bool bypass_l1) {
  constexpr int N = Tri<M2>::N;
#pragma unroll
  for (int i = 0; i < N; i++) T[i] = 0.0;
  // every thread takes one triangle per round; a round is merged cooperatively (warp, then CTA)
  for (int q0 = 0; q0 < count; q0 += TPB) {
    const int q = q0 + threadIdx.x;
    double S[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
      const double* src = tris + (long long)q * N + i;
      S[i] = (q < count) ? (bypass_l1 ? __ldcv(src) : *src) : 0.0;
    }
    if (q0 == 0) {
#pragma unroll
      for (int i = 0; i < N; i++) T[i] = S[i];
    } else {
      fold_tri<M2>(T, S);
    }
  }
  cta_merge_tri<M2, TPB / 32>(T, scratch);
}

#ifdef QRK_ROOT_TRACE
__device__ long long g_root_trace[16];
#define QRK_ROOT_CLK(i) do { if (threadIdx.x == 0) g_root_trace[i] = clock64(); } while (0)
#else
#define QRK_ROOT_CLK(i) do { } while (0)
#endif

template <int M2, int TPB, bool XCHG>
__global__ void __launch_bounds__(TPB)
angular_root_kernel(const double* __restrict__ tris, int count, int tris_ld, int mode, double* __restrict__ out_tri,
                    double* __restrict__ root, int* __restrict__ root_i, int keep_rhs_only, int* __restrict__ perm_tail,
                    int m1, AngularXchg xc) {
  using TR = Tri<M2>;
  constexpr int N = TR::N;
  __shared__ double scratch[(TPB / 32) * N];
  asm volatile("griddepcontrol.launch_dependents;");   // the back-substitution grid may be scheduled now (it waits for this grid)
  QRK_ROOT_CLK(0);
  // Every thread takes up to NT triangles (triangle q = r * TPB + t: a warp's loads of a component-major list are contiguous),
  // each warp merges its 32 * NT triangles cooperatively, warp 0 merges the warps' results: two dependent warp merges for up
  // to TPB * NT triangles, executed by TPB / 32 warps (few warps: the merges queue on the SM's shuffle port, see
  // warp_merge_tri_multi).  More triangles than that are folded serially into the first one beforehand.
  constexpr int NT = MergeFan<M2>::NT;
  double T[N];
  {
    double Tm[NT][N];
#pragma unroll
    for (int r = 0; r < NT; r++) {
      const int q = r * TPB + (int)threadIdx.x;
      const double* src = tris_ld ? tris + q : tris + (size_t)q * N;
      const size_t step = tris_ld ? (size_t)tris_ld : 1;
#pragma unroll
      for (int i = 0; i < N; i++) Tm[r][i] = (q < count) ? src[i * step] : 0.0;
    }
    for (int q0 = NT * TPB; q0 < count; q0 += TPB) {
      const int q = q0 + (int)threadIdx.x;
      const double* src = tris_ld ? tris + q : tris + (size_t)q * N;
      const size_t step = tris_ld ? (size_t)tris_ld : 1;
      double S[N];
#pragma unroll
      for (int i = 0; i < N; i++) S[i] = (q < count) ? src[i * step] : 0.0;
      fold_tri<M2>(Tm[0], S);
    }
    QRK_ROOT_CLK(1);
    warp_merge_tri_multi<M2, NT>(Tm);
#pragma unroll
    for (int i = 0; i < N; i++) T[i] = Tm[0][i];
  }
  if (TPB > 32) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < N; i++) scratch[warp * N + i] = T[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int i = 0; i < N; i++) T[i] = (lane < TPB / 32) ? scratch[lane * N + i] : 0.0;
      warp_merge_tri<M2>(T);
    }
    __syncthreads();            // scratch is reused by the exchange path
  }
  QRK_ROOT_CLK(2);
  bool poisoned = false;
  if constexpr (XCHG) {                          // mode 2 (a separate instantiation: the single-GPU root keeps its code)
    // ---- this GPU's triangle -> every rank's buffer (own included), then the flag; wait for all G flags
    const unsigned long long seq = *xc.seq + 1;               // the same on every rank: every rank runs the same sequence of steps
    const int G = xc.world, par = (int)(seq & 1ull);
    if (threadIdx.x == 0) {
      for (int g = 0; g < G; g++) {
        double* dst = xc.peers[g] + ((size_t)par * G + xc.rank) * N;
#pragma unroll
        for (int i = 0; i < N; i++) __stcg(dst + i, T[i]);
      }
      __threadfence_system();
      for (int g = 0; g < G; g++) {
        unsigned long long* flags = reinterpret_cast<unsigned long long*>(xc.peers[g] + (size_t)2 * G * N);
        *reinterpret_cast<volatile unsigned long long*>(flags + (size_t)par * G + xc.rank) = seq;
      }
    }
    if (threadIdx.x < G) {
      const volatile unsigned long long* mine =
          reinterpret_cast<const volatile unsigned long long*>(xc.peers[xc.rank] + (size_t)2 * G * N) + (size_t)par * G + threadIdx.x;
      const unsigned long long t0 = global_timer_ns();
      while (*mine != seq) {
        if (global_timer_ns() - t0 > xc.timeout_ns) { atomicExch(xc.err, 1); break; }
      }
      __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) *xc.seq = seq;
    const bool broken = *reinterpret_cast<volatile int*>(xc.err) != 0;   // this step or an earlier one lost a peer
    // ---- the G triangles in rank order: the same data, the same order, the same arithmetic on every rank
    if (G <= 32) {                            // one triangle per lane of warp 0: a single cooperative warp merge
      const double* src = xc.peers[xc.rank] + (size_t)par * G * N;
      if (threadIdx.x < 32) {
#pragma unroll
        for (int i = 0; i < N; i++) T[i] = ((int)threadIdx.x < G) ? __ldcv(src + (size_t)threadIdx.x * N + i) : 0.0;
        warp_merge_tri<M2>(T);
      }
    } else {
      merge_triangle_list<M2, TPB>(xc.peers[xc.rank] + (size_t)par * G * N, G, T, scratch, true);
    }
    if (broken) {
      const double poison = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
      for (int i = 0; i < N; i++) T[i] = poison;
      poisoned = true;
    }
  }
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
  QRK_ROOT_CLK(3);
  BlockQR<M2, M2, true, true>::run(a, tau, inv_diag, perm, z);
  QRK_ROOT_CLK(4);
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
      perm_tail[j] = m1 + perm[j];     // m_outputPerm_c(m1 + j) = m1 + P2(j) (BlockAngularSparseQR.h:501-503)
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
  if (poisoned) {       // a NaN triangle has rank 0 under the comparison above, which would zero y: keep the poison visible in x2 / x1
#pragma unroll
    for (int j = 0; j < M2; j++) y[j] = __longlong_as_double(0x7ff8000000000000ll);
  }
#pragma unroll
  for (int j = 0; j < M2; j++) root[M2 * M2 + M2 + j] = y[j];
#pragma unroll
  for (int j = 0; j < M2; j++) {
    root[M2 * M2 + 2 * M2 + perm[j]] = y[j];   // x2[P2[j]] = y_j (dest = P_c * y, :222)
  }
  QRK_ROOT_CLK(5);
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
  // Launched as a programmatic dependent of the root kernel (which triggers at its start): the CTAs that fit on the device
  // become resident while the root runs.  packed / atop / y1 are outputs of K1, which completed before the root started, so
  // they are fetched BEFORE the dependency wait and overlap the root's serial chain; only x2 (the root's output) is read
  // after it.  (K1 wrote them with evict_last: they are L2 hits.)
  const long long tile0 = (long long)blockIdx.x * TPB;
  const int count = (int)((nb - tile0 < TPB) ? (nb - tile0) : TPB);
#ifdef QRK_ANG_K3_LATE
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  stage_in_async<R * C, SA, TPB>(sA, packed + tile0 * (R * C), count);
  cp_async_commit();
  const long long blk = tile0 + t;
  const bool live = t < count;
  double y[R], at[M2][C];
#pragma unroll
  for (int k = 0; k < C; k++) y[k] = live ? __ldg(y1 + blk * C + k) : 0.0;
#pragma unroll
  for (int j = 0; j < M2; j++) {
#pragma unroll
    for (int k = 0; k < C; k++) at[j][k] = live ? __ldg(atop + (long long)j * m1 + blk * C + k) : 0.0;
  }
  cp_async_wait<0>();
  __syncthreads();
  double a[R * C];
  if (live) load_group<R * C>(a, sA + t * SA);
#ifndef QRK_ANG_K3_LATE
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  // x2 is the root kernel's output: loads pinned behind the wait (volatile asm; a plain load through the __restrict__ const
  // pointer may be hoisted above it), through L1 so that the 5 words cost one L2 request per SM rather than one per warp
  double x2[M2];
#pragma unroll
  for (int j = 0; j < M2; j++) asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(x2[j]) : "l"(root + M2 * M2 + 2 * M2 + j) : "memory");
  if (blockIdx.x == 0 && t < M2) {
    double v;
    asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(root + M2 * M2 + 2 * M2 + t) : "memory");
    x[m1 + t] = v;
  }
  if (!live) return;
#pragma unroll
  for (int j = 0; j < M2; j++) {
#pragma unroll
    for (int k = 0; k < C; k++) y[k] = fma(-at[j][k], x2[j], y[k]);
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

// K3 for 2 x 1 blocks: x1_p = (y1_p - Atop_p x2) / R11_p.  A programmatic dependent of the root: its CTAs become resident
// while the root's serial chain runs, and everything they need except x2 is an output of K1 (complete before the root
// started; L2 hits thanks to K1's evict_last stores).  One wave of CTAs therefore pulls its points ON CHIP before the
// dependency wait -- PR points per thread into registers and PS more into private shared-memory slots (cp.async; each
// thread reads back only what it copied, so no barrier) -- which holds ~90 % of a 1M-point problem (registers alone: 45 %);
// behind the wait only the 5-word x2 fetch, 6 FMAs a point and the store of x1 remain.  Points beyond the wave's capacity are
// swept afterwards.  (The one-point-per-thread kernel was bounded by the CTA launch rate: 7813 CTAs.)
template <int M2, int TPB, int PR, int PS>
__global__ void __launch_bounds__(TPB)
angular_backsolve_direct_kernel(const double* __restrict__ packed, const double* __restrict__ atop, const double* __restrict__ y1,
                                const double* root, double* __restrict__ x, long long nb64) {
  constexpr int NW = M2 + 2;                                  // words a point needs: R11, y1, Atop row
  extern __shared__ __align__(16) double slots[];             // [PS][NW][TPB]
  const unsigned nb = (unsigned)nb64, t = threadIdx.x;
  constexpr unsigned BATCH = TPB * (PR + PS);
  const unsigned base = blockIdx.x * BATCH + t;
  // shared-memory points of the first batch
#pragma unroll
  for (int i = 0; i < PS; i++) {
    const unsigned p = base + (PR + i) * TPB;
    if (p < nb) {
      cp_async8(slots + (i * NW + 0) * TPB + t, packed + 2 * (size_t)p);
      cp_async8(slots + (i * NW + 1) * TPB + t, y1 + p);
#pragma unroll
      for (int j = 0; j < M2; j++) cp_async8(slots + (i * NW + 2 + j) * TPB + t, atop + ((size_t)j * nb + p));
    }
  }
  cp_async_commit();
  // register points of the first batch
  double r11[PR], y[PR], at[PR][M2];
#pragma unroll
  for (int i = 0; i < PR; i++) {
    const unsigned p = base + i * TPB;
    const bool live = p < nb;
    r11[i] = live ? __ldg(packed + 2 * (size_t)p) : 1.0;
    y[i] = live ? __ldg(y1 + p) : 0.0;
#pragma unroll
    for (int j = 0; j < M2; j++) at[i][j] = live ? __ldg(atop + ((size_t)j * nb + p)) : 0.0;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // x2 is the root kernel's output: loads pinned behind the wait (volatile asm), through L1 (one L2 request per SM)
  double x2[M2];
#pragma unroll
  for (int j = 0; j < M2; j++) asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(x2[j]) : "l"(root + M2 * M2 + 2 * M2 + j) : "memory");
  if (blockIdx.x == 0 && t < M2) {
    double v;
    asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(root + M2 * M2 + 2 * M2 + t) : "memory");
    x[nb + t] = v;
  }
#pragma unroll
  for (int i = 0; i < PR; i++) {
    const unsigned p = base + i * TPB;
    if (p < nb) {
      double s = y[i];
#pragma unroll
      for (int j = 0; j < M2; j++) s = fma(-at[i][j], x2[j], s);
      x[p] = s * fast_rcp(r11[i]);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < PS; i++) {
    const unsigned p = base + (PR + i) * TPB;
    if (p < nb) {
      double s = slots[(i * NW + 1) * TPB + t];
#pragma unroll
      for (int j = 0; j < M2; j++) s = fma(-slots[(i * NW + 2 + j) * TPB + t], x2[j], s);
      x[p] = s * fast_rcp(slots[(i * NW + 0) * TPB + t]);
    }
  }
  // points beyond the first wave's capacity (grid capped at the resident CTAs): plain sweep
  for (unsigned q0 = (gridDim.x + blockIdx.x) * BATCH; q0 < nb; q0 += gridDim.x * BATCH) {
#pragma unroll
    for (int i = 0; i < PR + PS; i++) {
      const unsigned p = q0 + i * TPB + t;
      if (p < nb) {
        double s = __ldg(y1 + p);
#pragma unroll
        for (int j = 0; j < M2; j++) s = fma(-__ldg(atop + ((size_t)j * nb + p)), x2[j], s);
        x[p] = s * fast_rcp(__ldg(packed + 2 * (size_t)p));
      }
    }
  }
}

}  // namespace qrk
