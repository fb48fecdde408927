// export.cuh — matrixR() / matrixQ() in the reference's sparse formats, built on the device from the
// compact factors.  Index rules: reference src/QRKit/BlockDiagonalSparseQR.h:455-479 (FullQ),
// :483-500 (BlockDiagonalQ), identity tail :530-533; compressed layout as Eigen's setFromTriplets /
// insertBack+finalize produce it (inner indices ascending inside each outer vector).
// Not on the hot path: solve / apply never materialise Q or the sparse R.
#pragma once
#include "bd_generic.cuh"

namespace qrk {

// R: one thread per block. CSC n_rows x n_cols; column base_col+k holds rows (base_col|base_row)+j, j<=k.
__global__ void export_r_kernel(BlockIndex bi, const long long* __restrict__ eoff, long long nb,
                                const double* __restrict__ packed, int full_q, int* __restrict__ outer,
                                int* __restrict__ inner, double* __restrict__ vals) {
  const long long blk = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (blk >= nb) return;
  int r, c;
  long long vo, ro, co;
  bi.get(blk, r, c, vo, ro, co);
  const long long e0 = eoff ? eoff[blk] : blk * ((long long)c * (c + 1) / 2);
  const long long rbase = full_q ? co : ro;
  for (int k = 0; k < c; k++) {
    const long long p0 = e0 + (long long)k * (k + 1) / 2;
    outer[co + k] = (int)p0;
    for (int j = 0; j <= k; j++) {
      inner[p0 + j] = (int)(rbase + j);
      vals[p0 + j] = packed[vo + (long long)k * r + j];
    }
  }
}

__global__ void fill_outer_tail_kernel(int* outer, long long from, long long to_inclusive, long long base, int step) {
  for (long long i = from + blockIdx.x * (long long)blockDim.x + threadIdx.x; i <= to_inclusive;
       i += (long long)gridDim.x * blockDim.x)
    outer[i] = (int)(base + (i - from) * step);
}

// Q: one warp per (block, column kq of Q_i); Q_i e_kq = H_0 ... H_{c-1} e_kq.
template <int WPC>
__global__ void __launch_bounds__(32 * WPC)
export_q_kernel(BlockIndex bi, const long long* __restrict__ eoff, long long nb, const double* __restrict__ packed,
                const double* __restrict__ tau_in, long long n_cols, int full_q, int max_r, int* __restrict__ outer,
                int* __restrict__ inner, double* __restrict__ vals) {
  extern __shared__ __align__(16) double smem_q[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // items are enumerated block-major with max_r slots per block (slots >= r are idle)
  const long long item = (long long)blockIdx.x * WPC + warp;
  const long long blk = item / max_r;
  const int kq = (int)(item - blk * max_r);
  if (blk >= nb) return;
  int r, c;
  long long vo, ro, co;
  bi.get(blk, r, c, vo, ro, co);
  if (kq >= r) return;
  double* v = smem_q + (size_t)warp * max_r;
  const double* P = packed + vo;
  for (int i = lane; i < r; i += 32) v[i] = (i == kq) ? 1.0 : 0.0;
  __syncwarp();
  const int nv = r < c ? r : c;
  for (int k = nv - 1; k >= 0; --k) {
    const double* ck = P + (size_t)k * r;
    double dot = 0.0;
    for (int i = k + 1 + lane; i < r; i += 32) dot = fma(ck[i], v[i], dot);
    dot = warp_sum(dot) + v[k];
    const double tmp = tau_in[co + k] * dot;
    __syncwarp();
    for (int i = k + 1 + lane; i < r; i += 32) v[i] = fma(-ck[i], tmp, v[i]);
    if (lane == 0) v[k] -= tmp;
    __syncwarp();
  }
  const long long e0 = eoff ? eoff[blk] : blk * ((long long)r * r);
  const long long m1off = ro - co;
  const int col_index = full_q ? (int)(kq < c ? co + kq : n_cols + m1off + (kq - c)) : (int)(ro + kq);
  for (int j = lane; j < r; j += 32) {
    const long long p = e0 + (long long)j * r + kq;
    inner[p] = col_index;
    vals[p] = v[j];
    if (kq == 0) outer[ro + j] = (int)(e0 + (long long)j * r);
  }
}

__global__ void identity_tail_kernel(int* inner, double* vals, long long base, long long from_row, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    inner[base + i] = (int)(from_row + i);
    vals[base + i] = 1.0;
  }
}

// nnz_blocks = entries contributed by the blocks (the caller knows it from the block sizes).
inline cudaError_t launch_export_r(const BlockIndex& bi, const long long* d_eoff, long long nb, const double* packed,
                                   long long n_cols, long long sum_cols, long long nnz_blocks, int full_q, int* outer,
                                   int* inner, double* vals, cudaStream_t s) {
  if (nb > 0) export_r_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(bi, d_eoff, nb, packed, full_q, outer, inner, vals);
  // columns >= sum_cols are empty: their outer pointers, and outer[n_cols], all equal nnz
  fill_outer_tail_kernel<<<64, 256, 0, s>>>(outer, sum_cols, n_cols, nnz_blocks, 0);
  return cudaGetLastError();
}

inline cudaError_t launch_export_q(const BlockIndex& bi, const long long* d_eoff, long long nb, const double* packed,
                                   const double* tau, long long n_rows, long long n_cols, long long sum_rows,
                                   long long nnz_blocks, int full_q, int max_r, int* outer, int* inner, double* vals,
                                   cudaStream_t s) {
  constexpr int WPC = 4;
  if (nb > 0) {
    const long long items = nb * (long long)max_r;
    const size_t smem = (size_t)WPC * max_r * sizeof(double);
    auto kernel = export_q_kernel<WPC>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    kernel<<<(unsigned)((items + WPC - 1) / WPC), 32 * WPC, smem, s>>>(bi, d_eoff, nb, packed, tau, n_cols, full_q, max_r,
                                                                       outer, inner, vals);
  }
  // identity tail (BlockDiagonalSparseQR.h:530-533): row i >= sum_rows holds the single entry Q(i,i) = 1
  const long long tail = n_rows - sum_rows;
  if (tail > 0) identity_tail_kernel<<<64, 256, 0, s>>>(inner, vals, nnz_blocks, sum_rows, tail);
  fill_outer_tail_kernel<<<64, 256, 0, s>>>(outer, sum_rows, n_rows, nnz_blocks, 1);
  return cudaGetLastError();
}

}  // namespace qrk
