// banded_generic.cuh — the windowed recurrence of BandedBlockedSparseQR::factorize for ANY chain of dense windows
// (src/QRKit/BandedBlockedSparseQR.h:443-519, window step :494-507; Q products :640-675 / SparseBlockYTY.h:102-139).
//
// The templated kernels of banded.cuh cover six slab shapes at full speed (parallel groups + chase).  This file is the general
// path behind the same handle kind: block sizes, overlaps and column steps are run-time data, read from a window table
// {idxRow, idxCol, numRows, numCols} — the output of block detection on a general sparse matrix (qrk_detect_blocks) or one
// window per slab for slab shapes that are not instantiated.  It is the reference's own sequential schedule: one window after
// the other, the rows of R that are not yet final handed on to the next window; ONE CTA, the window in shared memory.
// It is latency bound by construction (a dependent chain of small Householder steps) and makes no attempt to be fast; what it
// adds is generality and an exact n x n Q:
//
//   window i works on  [carried rows (c_i) ; its own numRows_i matrix rows]  x  numCols_i columns (+ the right-hand side)
//   steps_i = min(rows, cols) Householder steps (Eigen makeHouseholder)
//   rows [0, solved_i)          final rows idxCol_i + r of R,  solved_i = idxCol_{i+1} - idxCol_i  (all steps_i for the last window)
//   rows [solved_i, steps_i)    carried into window i+1 (they overlap its first columns)
//   rows [steps_i, rows)        annihilated: the complement, appended to the output of Q^T v in window order
//
// so Q^T v = [ thin part, n_cols values ; complement, n_rows - n_cols values ] with Q orthogonal n_rows x n_rows, and Q v is its
// inverse (the same reflectors backwards).  R is unique up to row signs for a fixed column order, so it equals the reference's
// whatever the window blocking; the stored pattern of matrixR() still follows the reference's merged windows (capi.cu).
#pragma once
#include "banded_dispatch.hpp"
#include "bd_generic.cuh"

namespace qrk {

constexpr int kGenThreads = 256;

__device__ __forceinline__ double gen_block_sum(double v, double* sred) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sred[warp] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < kGenThreads / 32; w++) s += sred[w];
  return s;
}

// Eigen makeHouseholder on column k of the window W (ld = leading dimension, rows valid rows), by the whole CTA.
// Leaves beta on the diagonal, the essential part below, tau in *tau_out (shared), and returns tau.
__device__ __forceinline__ double gen_make_reflector(double* W, int ld, int rows, int k, double* sred, double* sscal) {
  double* col = W + (size_t)k * ld;
  double part = 0.0;
  for (int r = k + 1 + threadIdx.x; r < rows; r += kGenThreads) part = fma(col[r], col[r], part);
  const double tailSq = gen_block_sum(part, sred);
  const double c0 = col[k];
  double beta, tau, inv;
  if (tailSq <= DBL_MIN) { tau = 0.0; beta = c0; inv = 0.0; }
  else {
    beta = sqrt(fma(c0, c0, tailSq));
    if (c0 >= 0.0) beta = -beta;
    inv = 1.0 / (c0 - beta);
    tau = (beta - c0) / beta;
  }
  __syncthreads();                                   // every thread has read col[k]
  for (int r = k + 1 + threadIdx.x; r < rows; r += kGenThreads) col[r] *= inv;
  if (threadIdx.x == 0) { col[k] = beta; sscal[0] = tau; }
  __syncthreads();
  return tau;
}

// H_k applied to the columns [j0, j1) of W: one warp per column, rows across the lanes
__device__ __forceinline__ void gen_apply_reflector(double* W, int ld, int rows, int k, double tau, int j0, int j1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* v = W + (size_t)k * ld;
  for (int j = j0 + warp; j < j1; j += kGenThreads / 32) {
    double* cj = W + (size_t)j * ld;
    double dot = 0.0;
    for (int r = k + 1 + lane; r < rows; r += 32) dot = fma(v[r], cj[r], dot);
    dot = warp_sum(dot) + cj[k];
    const double w = tau * dot;
    for (int r = k + 1 + lane; r < rows; r += 32) cj[r] = fma(-v[r], w, cj[r]);
    __syncwarp();
    if (lane == 0) cj[k] -= w;
  }
}

// ---- factorize (+ fused Q^T b when a.b != nullptr): the window chain, sequentially ----------------------------------------
// dynamic shared memory: W[(max_cols + 1) * ldw] + carry[(max_cols + 1) * max_cols'] ... see gen_smem_bytes
__global__ void __launch_bounds__(kGenThreads) banded_generic_factor_kernel(GenArgs a) {
  extern __shared__ __align__(16) double gsm[];
  __shared__ double sred[kGenThreads / 32];
  __shared__ double sscal[2];
  const int ldw = a.max_rows | 1;                     // odd leading dimension: rows of a column spread over the banks
  double* W = gsm;                                    // (max_cols + 1) columns; the last used column is the right-hand side
  double* Cb = W + (size_t)(a.max_cols + 1) * ldw;    // carried rows: max_cols rows x (max_cols + 1) columns, leading dimension ldc
  const int ldc = a.max_cols | 1;
  const bool rhs = a.b != nullptr;
  int prev_solved = 0, prev_cols = 0;
  for (int i = 0; i < a.nwin; i++) {
    const GenWindow w = a.win[i];
    const int rows = w.carry + w.nrows, cols = w.ncols, ncx = cols + (rhs ? 1 : 0);
    // assemble: carried rows on top (their columns [prev_solved, prev_cols) become columns [0, prev_cols - prev_solved)), zeros elsewhere
    for (int e = threadIdx.x; e < ncx * rows; e += kGenThreads) {
      const int j = e / rows, r = e - j * rows;
      double v = 0.0;
      if (r < w.carry) {
        if (j == cols) v = Cb[(size_t)a.max_cols * ldc + r];
        else if (j < prev_cols - prev_solved) v = Cb[(size_t)j * ldc + r];
      } else {
        const int rr = r - w.carry;
        if (j == cols) v = a.b[w.row0 + rr];
        else if (j < w.ncols_in) v = a.A_in[w.voff + (long long)j * w.nrows + rr];
      }
      W[(size_t)j * ldw + r] = v;
    }
    __syncthreads();
    for (int k = 0; k < w.steps; k++) {
      const double tau = gen_make_reflector(W, ldw, rows, k, sred, sscal);
      if (threadIdx.x == 0) a.tau[w.toff + k] = tau;
      gen_apply_reflector(W, ldw, rows, k, tau, k + 1, ncx);
      __syncthreads();
    }
    // packed window for the Q products; the finished rows' right-hand side; the carried rows; the complement
    for (int e = threadIdx.x; e < cols * rows; e += kGenThreads) {
      const int j = e / rows, r = e - j * rows;
      a.packed[w.poff + (long long)j * rows + r] = W[(size_t)j * ldw + r];
    }
    if (rhs) {
      for (int r = threadIdx.x; r < w.solved; r += kGenThreads) a.y[w.col0 + r] = W[(size_t)cols * ldw + r];
      if (a.comp) for (int r = w.steps + threadIdx.x; r < rows; r += kGenThreads) a.comp[w.coff + r - w.steps] = W[(size_t)cols * ldw + r];
    }
    const int ncarry = w.steps - w.solved;
    for (int e = threadIdx.x; e < (cols - w.solved + 1) * ncarry; e += kGenThreads) {
      const int jj = e / ncarry, r = e - jj * ncarry;              // jj = cols - solved: the right-hand side
      const bool is_rhs = jj == cols - w.solved;
      const double v = is_rhs ? (rhs ? W[(size_t)cols * ldw + w.solved + r] : 0.0)
                              : ((w.solved + jj >= w.solved + r) ? W[(size_t)(w.solved + jj) * ldw + w.solved + r] : 0.0);   // upper trapezoid only
      Cb[(size_t)(is_rhs ? a.max_cols : jj) * ldc + r] = v;
    }
    prev_solved = w.solved; prev_cols = cols;
    __syncthreads();
  }
}

// ---- Q^T v / Q v on the stored windows (one vector per CTA: blockIdx.x = column) ----------------------------------------------
// transpose: in = v (n_rows, leading dimension ldin) -> y (thin, ldy) and comp (ldcomp, may be null)
// forward  : y (thin) and comp (null = zero complement) -> out (n_rows)
template <bool TRANSPOSE>
__global__ void __launch_bounds__(kGenThreads) banded_generic_apply_kernel(GenArgs a, const double* in, long long ldin, double* thin, long long ldthin,
                                                                           double* comp, long long ldcomp, double* out, long long ldout) {
  extern __shared__ __align__(16) double gsm[];
  __shared__ double sred[kGenThreads / 32];
  double* wv = gsm;                                   // max_rows
  double* carry = gsm + a.max_rows;                   // max_cols
  const long long col = blockIdx.x;
  if (in) in += col * ldin;
  if (thin) thin += col * ldthin;
  if (comp) comp += col * ldcomp;
  if (out) out += col * ldout;
  if (TRANSPOSE) {
    for (int i = 0; i < a.nwin; i++) {
      const GenWindow w = a.win[i];
      const int rows = w.carry + w.nrows;
      for (int r = threadIdx.x; r < rows; r += kGenThreads) wv[r] = (r < w.carry) ? carry[r] : in[w.row0 + r - w.carry];
      __syncthreads();
      const double* P = a.packed + w.poff;
      for (int k = 0; k < w.steps; k++) {
        const double* v = P + (long long)k * rows;
        double dot = 0.0;
        for (int r = k + 1 + threadIdx.x; r < rows; r += kGenThreads) dot = fma(v[r], wv[r], dot);
        dot = gen_block_sum(dot, sred) + wv[k];
        const double t = a.tau[w.toff + k] * dot;
        __syncthreads();
        for (int r = k + 1 + threadIdx.x; r < rows; r += kGenThreads) wv[r] = fma(-v[r], t, wv[r]);
        if (threadIdx.x == 0) wv[k] -= t;
        __syncthreads();
      }
      for (int r = threadIdx.x; r < w.solved; r += kGenThreads) thin[w.col0 + r] = wv[r];
      for (int r = w.solved + threadIdx.x; r < w.steps; r += kGenThreads) carry[r - w.solved] = wv[r];
      if (comp) for (int r = w.steps + threadIdx.x; r < rows; r += kGenThreads) comp[w.coff + r - w.steps] = wv[r];
      __syncthreads();
    }
  } else {
    for (int i = a.nwin - 1; i >= 0; i--) {
      const GenWindow w = a.win[i];
      const int rows = w.carry + w.nrows;
      for (int r = threadIdx.x; r < rows; r += kGenThreads)
        wv[r] = (r < w.solved) ? thin[w.col0 + r] : (r < w.steps) ? carry[r - w.solved] : (comp ? comp[w.coff + r - w.steps] : 0.0);
      __syncthreads();
      const double* P = a.packed + w.poff;
      for (int k = w.steps - 1; k >= 0; k--) {
        const double* v = P + (long long)k * rows;
        double dot = 0.0;
        for (int r = k + 1 + threadIdx.x; r < rows; r += kGenThreads) dot = fma(v[r], wv[r], dot);
        dot = gen_block_sum(dot, sred) + wv[k];
        const double t = a.tau[w.toff + k] * dot;
        __syncthreads();
        for (int r = k + 1 + threadIdx.x; r < rows; r += kGenThreads) wv[r] = fma(-v[r], t, wv[r]);
        if (threadIdx.x == 0) wv[k] -= t;
        __syncthreads();
      }
      for (int r = threadIdx.x; r < rows; r += kGenThreads) {
        if (r < w.carry) carry[r] = wv[r];
        else out[w.row0 + r - w.carry] = wv[r];
      }
      __syncthreads();
    }
  }
}

// ---- x = R^-1 y: rows from the bottom, one CTA; row g lives in its window's packed block (row g - col0, columns >= that) ----
__global__ void __launch_bounds__(kGenThreads) banded_generic_backsolve_kernel(GenArgs a, const double* y, double* x) {
  __shared__ double sred[kGenThreads / 32];
  for (int i = a.nwin - 1; i >= 0; i--) {
    const GenWindow w = a.win[i];
    const int rows = w.carry + w.nrows;
    const double* P = a.packed + w.poff;
    for (int r = w.solved - 1; r >= 0; r--) {
      const long long g = w.col0 + r;
      double part = 0.0;
      for (int j = r + 1 + threadIdx.x; j < w.ncols; j += kGenThreads) part = fma(P[(long long)j * rows + r], x[w.col0 + j], part);
      const double s = gen_block_sum(part, sred);
      if (threadIdx.x == 0) x[g] = (y[g] - s) / P[(long long)r * rows + r];
      __syncthreads();
    }
  }
}

// ---- values of matrixR() for a given compressed pattern: entry (g, j) = R(g, j) from the window that finalised row g ----
__global__ void banded_generic_export_r_kernel(GenArgs a, const int* __restrict__ win_col0 /* nwin + 1 */, const int* __restrict__ outer,
                                               const int* __restrict__ inner, double* __restrict__ vals) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < a.n_cols; j += (long long)gridDim.x * blockDim.x) {
    for (int p = outer[j]; p < outer[j + 1]; p++) {
      const long long g = inner[p];
      double v = 0.0;
      if (g <= j) {
        int lo = 0, hi = a.nwin - 1;                  // last window with col0 <= g
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (win_col0[mid] <= g) lo = mid; else hi = mid - 1; }
        const GenWindow w = a.win[lo];
        const int r = (int)(g - w.col0), c = (int)(j - w.col0);
        if (r < w.solved && c < w.ncols) v = a.packed[w.poff + (long long)c * (w.carry + w.nrows) + r];
      }
      vals[p] = v;
    }
  }
}

inline size_t gen_smem_factor(int max_rows, int max_cols) {
  return ((size_t)(max_cols + 1) * (max_rows | 1) + (size_t)(max_cols + 1) * (max_cols | 1)) * sizeof(double);
}
inline size_t gen_smem_apply(int max_rows, int max_cols) { return (size_t)(max_rows + max_cols) * sizeof(double); }

}  // namespace qrk
