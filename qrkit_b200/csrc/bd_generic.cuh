// bd_generic.cuh — batched FP64 Householder QR of diagonal blocks of ANY size: one team of W warps
// (W = 1: a warp, W > 1: the whole CTA) per block, the block resident in shared memory.
//
// Same reference path as bd_small.cuh (BlockDiagonalSparseQR::factorize, reference
// src/QRKit/BlockDiagonalSparseQR.h:432-526, and _solve_impl :258-280) for blocks that do not fit in
// one thread's registers or whose size is not known at compile time (SparseBlockDiagonal<MatrixXd>,
// BASELINE config 5: 32x16 ... 128x64 mixed).  The block is one contiguous r*c*8-byte range of the
// block-COO array, so it is fetched with ONE bulk asynchronous copy (cp.async.bulk, the 1-D TMA path,
// SASS UBLKCP) signalled on an mbarrier, factorised in place in shared memory (reflector norms and
// the v^T A dot products are warp-shuffle reductions), and written back with one bulk store.
#pragma once
#include "common.cuh"

namespace qrk {

// Block index of the block-COO storage (device arrays; all null => uniform ur x uc blocks).
struct BlockIndex {
  const int* rows;
  const int* cols;
  const long long* voff;   // offset of block i in values[]
  const long long* roff;   // base_row(i)
  const long long* coff;   // base_col(i)
  int ur, uc;
  __device__ __forceinline__ void get(long long i, int& r, int& c, long long& vo, long long& ro, long long& co) const {
    if (rows) { r = rows[i]; c = cols[i]; vo = voff[i]; ro = roff[i]; co = coff[i]; }
    else { r = ur; c = uc; vo = i * (long long)ur * uc; ro = i * ur; co = i * uc; }
  }
};

// ---- mbarrier / bulk-copy wrappers (PTX ISA: cp.async.bulk, mbarrier) ---------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem)),
               "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem, const void* smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem), "r"(smem_u32(smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int W>
__device__ __forceinline__ void team_sync() {
  if (W == 1) __syncwarp();
  else __syncthreads();
}

// Shared-memory footprint (bytes) of a team for an r x c block.
__host__ __device__ inline size_t generic_smem_bytes(int r, int c) {
  size_t d = (size_t)r * c + r /*rhs*/ + 3 * (size_t)c /*tau, upd, dir*/;
  d = (d + 1) & ~(size_t)1;
  return d * 8 + (size_t)c * 4 /*perm*/ + 32 /*mbarrier + scalars*/ + 16;
}

// ---------------------------------------------------------------------------------------------
// Factorize (+ fused solve) kernel.  grid.x = number of blocks in this launch; ids (optional) maps
// launch-local index -> block number (size classes are launched separately so that the dynamic
// shared memory, and with it the occupancy, matches the class).
// ---------------------------------------------------------------------------------------------
template <int W, bool PIV, bool SOLVE>
__global__ void __launch_bounds__(32 * W)
bd_generic_factor_kernel(BlockIndex bi, const int* __restrict__ ids, const double* A_in, double* packed,
                         double* __restrict__ tau_out, int* __restrict__ perm_out, const double* __restrict__ b,
                         double* __restrict__ x) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const long long blk = ids ? ids[blockIdx.x] : blockIdx.x;
  int r, c;
  long long vo, ro, co;
  bi.get(blk, r, c, vo, ro, co);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int T = 32 * W;

  double* sA = reinterpret_cast<double*>(smem_raw);
  double* sRhs = sA + (size_t)r * c;
  double* sTau = sRhs + r;
  double* sUpd = sTau + c;
  double* sDir = sUpd + c;
  size_t dcount = ((size_t)r * c + r + 3 * (size_t)c + 1) & ~(size_t)1;
  int* sPerm = reinterpret_cast<int*>(sA + dcount);
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sPerm) + (((size_t)c * 4 + 15) & ~(size_t)15));
  int* sBig = reinterpret_cast<int*>(bar + 1);

  const double* gA = A_in + vo;
  const uint32_t bytes = (uint32_t)((size_t)r * c * 8);
  const bool bulk_ok = ((reinterpret_cast<uintptr_t>(gA) & 15) == 0) && ((bytes & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(packed + vo) & 15) == 0);
  if (bulk_ok) {
    if (tid == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    team_sync<W>();
    if (tid == 0) {
      mbar_expect_tx(bar, bytes);
      // bulk copies are limited to < 2^20 bytes per instruction by our own chunking for safety
      uint32_t done = 0;
      while (done < bytes) {
        const uint32_t chunk = (bytes - done > 65536u) ? 65536u : (bytes - done);
        bulk_g2s(reinterpret_cast<unsigned char*>(sA) + done, reinterpret_cast<const unsigned char*>(gA) + done, chunk, bar);
        done += chunk;
      }
    }
  } else {
    for (int i = tid; i < r * c; i += T) sA[i] = gA[i];
  }
  if (SOLVE) for (int i = tid; i < r; i += T) sRhs[i] = b[ro + i];
  for (int j = tid; j < c; j += T) { sPerm[j] = j; sTau[j] = 0.0; }
  if (bulk_ok) mbar_wait(bar, 0);
  team_sync<W>();

  const int nv = r < c ? r : c;
  const int ncol_ext = SOLVE ? c + 1 : c;   // the rhs rides along as column c

  if (PIV) {
    for (int j = warp; j < c; j += W) {
      double s = 0.0;
      for (int i = lane; i < r; i += 32) { const double v = sA[(size_t)j * r + i]; s = fma(v, v, s); }
      s = warp_sum(s);
      if (lane == 0) { sUpd[j] = sDir[j] = sqrt(s); }
    }
    team_sync<W>();
  }

  for (int k = 0; k < nv; k++) {
    if (PIV) {
      if (warp == 0) {
        // first maximum of upd[k..c-1]
        double bv = -1.0;
        int bj = 0x7fffffff;
        for (int j = k + lane; j < c; j += 32) {
          const double u = sUpd[j];
          if (u > bv) { bv = u; bj = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
          if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
        }
        if (lane == 0) {
          *sBig = bj;
          if (bj != k) {
            double t = sUpd[k]; sUpd[k] = sUpd[bj]; sUpd[bj] = t;
            t = sDir[k]; sDir[k] = sDir[bj]; sDir[bj] = t;
            const int p = sPerm[k]; sPerm[k] = sPerm[bj]; sPerm[bj] = p;
          }
        }
      }
      team_sync<W>();
      const int big = *sBig;
      if (big != k) {
        for (int i = tid; i < r; i += T) {
          const double t = sA[(size_t)k * r + i];
          sA[(size_t)k * r + i] = sA[(size_t)big * r + i];
          sA[(size_t)big * r + i] = t;
        }
      }
      team_sync<W>();
    }
    // (a) reflector of column k, computed redundantly by every warp (read-only on column k)
    const double* ck = sA + (size_t)k * r;
    double tailSq = 0.0;
    for (int i = k + 1 + lane; i < r; i += 32) { const double v = ck[i]; tailSq = fma(v, v, tailSq); }
    tailSq = warp_sum(tailSq);
    const double c0 = ck[k];
    const bool degenerate = (k + 1 >= r) || (tailSq <= DBL_MIN);
    double beta, inv, tau;                      // Eigen makeHouseholder (one short dependent chain: common.cuh)
    householder_scalars(c0, tailSq, degenerate, beta, inv, tau);
    // (b) trailing columns (and the rhs): col -= tau * v * (v^T col), v = [1; inv*ck[k+1:]].  Warp w owns the columns
    //     k+1+w, k+1+w+W, ...; for blocks of up to 128 rows it takes them FOUR at a time with the reflector and the four
    //     columns in registers: the four warp reductions run side by side (one shuffle latency chain per batch instead of
    //     one per column) and lane q does the LAWN-176 norm downdate of the batch's q-th column.  Same per-lane summation
    //     order as the one-column form below, so the results are bit-identical.
    if (r <= 128) {
      constexpr int RPL = 4, NBAT = 4;
      double v[RPL];
#pragma unroll
      for (int m = 0; m < RPL; m++) { const int i = k + 1 + lane + 32 * m; v[m] = (i < r) ? ck[i] : 0.0; }
      for (int j0 = k + 1 + warp; j0 < ncol_ext; j0 += NBAT * W) {
        double cv[NBAT][RPL], dot[NBAT], piv[NBAT];
#pragma unroll
        for (int q = 0; q < NBAT; q++) {
          const int j = j0 + q * W;
          const double* cj = (j < c) ? sA + (size_t)j * r : sRhs;
          dot[q] = 0.0;
#pragma unroll
          for (int m = 0; m < RPL; m++) {
            const int i = k + 1 + lane + 32 * m;
            cv[q][m] = (j < ncol_ext && i < r) ? cj[i] : 0.0;
            dot[q] = fma(v[m], cv[q][m], dot[q]);
          }
          piv[q] = (j < ncol_ext) ? cj[k] : 0.0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int q = 0; q < NBAT; q++) dot[q] += __shfl_xor_sync(0xffffffffu, dot[q], o);
        }
        double newk[NBAT], nsq[NBAT];
#pragma unroll
        for (int q = 0; q < NBAT; q++) {
          const int j = j0 + q * W;
          double* cj = (j < c) ? sA + (size_t)j * r : sRhs;
          const double tmp = tau * fma(dot[q], inv, piv[q]);
          const double tinv = tmp * inv;
          nsq[q] = 0.0;
#pragma unroll
          for (int m = 0; m < RPL; m++) {
            const int i = k + 1 + lane + 32 * m;
            cv[q][m] = fma(-v[m], tinv, cv[q][m]);
            if (j < ncol_ext && i < r) cj[i] = cv[q][m];
            nsq[q] = fma(cv[q][m], cv[q][m], nsq[q]);
          }
          newk[q] = piv[q] - tmp;
          if (lane == 0 && j < ncol_ext) cj[k] = newk[q];
        }
        if (PIV) {
          // lane q: downdate of column j0 + q W; a column that must be recomputed gets its norm from the registers
          bool recompute = false;
          int jq = -1;
#pragma unroll
          for (int q = 0; q < NBAT; q++) if (lane == q) { jq = j0 + q * W; }
          double myk = 0.0;
#pragma unroll
          for (int q = 0; q < NBAT; q++) myk = (lane == q) ? newk[q] : myk;
          if (lane < NBAT && jq < c) {
            const double upd = sUpd[jq];
            if (upd != 0.0) {
              double t = fabs(myk) / upd;
              t = (1.0 + t) * (1.0 - t);
              t = t < 0.0 ? 0.0 : t;
              const double qq = upd / sDir[jq];
              const double t2 = t * (qq * qq);
              if (t2 <= 1.4901161193847656e-08) recompute = true;
              else sUpd[jq] = upd * sqrt(t);
            }
          }
          const unsigned need = __ballot_sync(0xffffffffu, recompute);
          if (need) {
#pragma unroll
            for (int q = 0; q < NBAT; q++) {
              if (need & (1u << q)) {
                const double su = warp_sum(nsq[q]);
                const int j = j0 + q * W;
                if (lane == 0) { sDir[j] = sqrt(su); sUpd[j] = sDir[j]; }
              }
            }
          }
        }
      }
    } else {
    for (int j = k + 1 + warp; j < ncol_ext; j += W) {
      double* cj = (j < c) ? sA + (size_t)j * r : sRhs;
      double dot = 0.0;
      for (int i = k + 1 + lane; i < r; i += 32) dot = fma(ck[i], cj[i], dot);
      dot = warp_sum(dot);
      dot = fma(dot, inv, cj[k]);
      const double tmp = tau * dot;
      const double tinv = tmp * inv;
      for (int i = k + 1 + lane; i < r; i += 32) cj[i] = fma(-ck[i], tinv, cj[i]);
      __syncwarp();
      if (lane == 0) cj[k] -= tmp;
      if (PIV && j < c) {
        // LAWN-176 downdate of column j's partial norm, by lane 0 of the owning warp
        double su = 0.0;
        bool recompute = false;
        if (lane == 0) {
          const double upd = sUpd[j];
          if (upd != 0.0) {
            double t = fabs(cj[k]) / upd;
            t = (1.0 + t) * (1.0 - t);
            t = t < 0.0 ? 0.0 : t;
            const double q = upd / sDir[j];
            const double t2 = t * (q * q);
            if (t2 <= 1.4901161193847656e-08) recompute = true;
            else sUpd[j] = upd * sqrt(t);
          }
        }
        recompute = __shfl_sync(0xffffffffu, (int)recompute, 0) != 0;
        if (recompute) {
          for (int i = k + 1 + lane; i < r; i += 32) su = fma(cj[i], cj[i], su);
          su = warp_sum(su);
          if (lane == 0) { sDir[j] = sqrt(su); sUpd[j] = sDir[j]; }
        }
      }
    }
    }
    team_sync<W>();
    // (c) store the essential part and beta (every reader of the raw column is past the barrier)
    if (warp == (k % W)) {
      double* wk = sA + (size_t)k * r;
      for (int i = k + 1 + lane; i < r; i += 32) wk[i] *= inv;
      if (lane == 0) { wk[k] = beta; sTau[k] = tau; }
    }
    // no barrier needed: the next iteration reads columns > k only, the epilogue is behind a barrier
  }
  team_sync<W>();

  if (SOLVE) {
    // back substitution on the c x c upper triangle, by warp 0
    if (warp == 0) {
      for (int j = c - 1; j >= 0; --j) {
        const double yj = sRhs[j] / sA[(size_t)j * r + j];
        __syncwarp();
        if (lane == 0) sRhs[j] = yj;
        for (int i = lane; i < j; i += 32) sRhs[i] = fma(-sA[(size_t)j * r + i], yj, sRhs[i]);
        __syncwarp();
      }
    }
    team_sync<W>();
    for (int j = tid; j < c; j += T) x[co + sPerm[j]] = sRhs[j];
  }
  for (int j = tid; j < c; j += T) {
    tau_out[co + j] = sTau[j];
    if (PIV) perm_out[co + j] = (int)co + sPerm[j];
  }
  if (bulk_ok) {
    fence_async_smem();     // generic-proxy writes to smem -> visible to the async proxy
    team_sync<W>();
    if (tid == 0) {
      uint32_t done = 0;
      while (done < bytes) {
        const uint32_t chunk = (bytes - done > 65536u) ? 65536u : (bytes - done);
        bulk_s2g(reinterpret_cast<unsigned char*>(packed + vo) + done, reinterpret_cast<unsigned char*>(sA) + done, chunk);
        done += chunk;
      }
      bulk_commit();
      bulk_wait_all();
    }
  } else {
    for (int i = tid; i < r * c; i += T) packed[vo + i] = sA[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Operations on an existing factorisation: one WARP per (block, rhs); reflectors are streamed from
// global memory (each read once per rhs), the vector lives in shared memory.
// op: OP_SOLVE / OP_APPLY_QT / OP_APPLY_Q as in bd_small.cuh; full_q selects the FullQ index layout.
// ---------------------------------------------------------------------------------------------
// column `col` of the packed block, rows lo < i < hi of this lane (i = lane + 32 m), zeros elsewhere
template <int RPL>
__device__ __forceinline__ void op_fetch_col(double (&dst)[RPL], const double* __restrict__ P, int r, int col, bool valid, int lo, int hi, int lane) {
#pragma unroll
  for (int m = 0; m < RPL; m++) {
    const int i = lane + 32 * m;
    dst[m] = (valid && i > lo && i < hi) ? __ldg(P + (size_t)col * r + i) : 0.0;
  }
}

template <int WPC>   // warps per CTA
__global__ void __launch_bounds__(32 * WPC)
bd_generic_op_kernel(BlockIndex bi, long long nb, const double* __restrict__ packed, const double* __restrict__ tau_in,
                     const int* __restrict__ perm, const double* __restrict__ B, long long ldb, double* __restrict__ X,
                     long long ldx, int nrhs, long long n_cols, int op, int full_q, int max_r) {
  extern __shared__ __align__(16) double smem_d[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long item = (long long)blockIdx.x * WPC + warp;
  if (item >= nb * nrhs) return;
  const long long blk = item % nb;
  const int rhs_i = (int)(item / nb);
  int r, c;
  long long vo, ro, co;
  bi.get(blk, r, c, vo, ro, co);
  double* v = smem_d + (size_t)warp * max_r;
  const double* Bc = B + (long long)rhs_i * ldb;
  double* Xc = X + (long long)rhs_i * ldx;
  const double* P = packed + vo;
  const long long m1off = ro - co;
  const int nv = r < c ? r : c;

  if (op == 2 /*OP_APPLY_Q*/ && full_q) {
    for (int i = lane; i < c; i += 32) v[i] = Bc[co + i];
    for (int i = lane; i < r - c; i += 32) v[c + i] = Bc[n_cols + m1off + i];
  } else {
    for (int i = lane; i < r; i += 32) v[i] = Bc[ro + i];
  }
  __syncwarp();
  // The reflector columns are streamed from HBM in chunks of KB columns held in registers, the next chunk in flight while the
  // current one is applied: the chain per reflector is then one warp reduction, not one memory round trip (blocks of up to
  // 32 * RPL rows; taller blocks fall back to direct loads).
  constexpr int KB = 4, RPL = 4;
  if (r <= 32 * RPL) {
    double cur[KB][RPL], nxt[KB][RPL];
#define QRK_OP_FETCH1(dst, kk0_, q)                                                               \
  { const int kkf = (kk0_) + q, kf = (op == 2) ? nv - 1 - kkf : kkf;                               \
    op_fetch_col<RPL>(dst[q], P, r, kkf < nv ? kf : 0, kkf < nv, kf, r, lane); }
#define QRK_OP_FETCH(dst, kk0_) QRK_OP_FETCH1(dst, kk0_, 0) QRK_OP_FETCH1(dst, kk0_, 1) QRK_OP_FETCH1(dst, kk0_, 2) QRK_OP_FETCH1(dst, kk0_, 3)
    static_assert(KB == 4, "QRK_OP_FETCH is written out for four columns");
    QRK_OP_FETCH(cur, 0)
    for (int kk0 = 0; kk0 < nv; kk0 += KB) {
      QRK_OP_FETCH(nxt, kk0 + KB)
#pragma unroll
      for (int q = 0; q < KB; q++) {
        const int kk = kk0 + q;
        if (kk < nv) {
          const int k = (op == 2) ? nv - 1 - kk : kk;
          const double tau = tau_in[co + k];
          double dot = 0.0;
#pragma unroll
          for (int m = 0; m < RPL; m++) { const int i = lane + 32 * m; if (i < r) dot = fma(cur[q][m], v[i], dot); }
          dot = warp_sum(dot) + v[k];
          const double tmp = tau * dot;
          __syncwarp();
#pragma unroll
          for (int m = 0; m < RPL; m++) { const int i = lane + 32 * m; if (i > k && i < r) v[i] = fma(-cur[q][m], tmp, v[i]); }
          if (lane == 0) v[k] -= tmp;
          __syncwarp();
        }
      }
#pragma unroll
      for (int q = 0; q < KB; q++)
#pragma unroll
        for (int m = 0; m < RPL; m++) cur[q][m] = nxt[q][m];
    }
  } else {
  for (int kk = 0; kk < nv; kk++) {
    const int k = (op == 2) ? nv - 1 - kk : kk;
    const double* ck = P + (size_t)k * r;
    const double tau = tau_in[co + k];
    double dot = 0.0;
    for (int i = k + 1 + lane; i < r; i += 32) dot = fma(ck[i], v[i], dot);
    dot = warp_sum(dot) + v[k];
    const double tmp = tau * dot;
    __syncwarp();
    for (int i = k + 1 + lane; i < r; i += 32) v[i] = fma(-ck[i], tmp, v[i]);
    if (lane == 0) v[k] -= tmp;
    __syncwarp();
  }
  }
  if (op == 0 /*OP_SOLVE*/) {
    if (c <= 32 * RPL) {
      double cur[KB][RPL], nxt[KB][RPL];
#define QRK_OP_FETCH_R1(dst, jj0_, q)                         /* columns c-1-jj0, c-2-jj0, ...: rows <= the diagonal */ \
  { const int jf = c - 1 - ((jj0_) + q);                                                          \
    op_fetch_col<RPL>(dst[q], P, r, jf >= 0 ? jf : 0, jf >= 0, -1, jf + 1, lane); }
#define QRK_OP_FETCH_R(dst, jj0_) QRK_OP_FETCH_R1(dst, jj0_, 0) QRK_OP_FETCH_R1(dst, jj0_, 1) QRK_OP_FETCH_R1(dst, jj0_, 2) QRK_OP_FETCH_R1(dst, jj0_, 3)
      QRK_OP_FETCH_R(cur, 0)
      for (int jj0 = 0; jj0 < c; jj0 += KB) {
        QRK_OP_FETCH_R(nxt, jj0 + KB)
#pragma unroll
        for (int q = 0; q < KB; q++) {
          const int j = c - 1 - (jj0 + q);
          if (j >= 0) {
            double dg = 0.0;                                         // R_jj sits in lane j % 32, slot j / 32: one broadcast per
#pragma unroll                                                       // slot (a register array must not be indexed by j)
            for (int m = 0; m < RPL; m++) {
              const double dm = __shfl_sync(0xffffffffu, cur[q][m], j & 31);
              dg = ((j >> 5) == m) ? dm : dg;
            }
            const double yj = v[j] / dg;
            __syncwarp();
#pragma unroll
            for (int m = 0; m < RPL; m++) { const int i = lane + 32 * m; if (i < j) v[i] = fma(-cur[q][m], yj, v[i]); }
            if (lane == 0) v[j] = yj;
            __syncwarp();
          }
        }
#pragma unroll
        for (int q = 0; q < KB; q++)
#pragma unroll
          for (int m = 0; m < RPL; m++) cur[q][m] = nxt[q][m];
      }
    } else {
    for (int j = c - 1; j >= 0; --j) {
      const double yj = v[j] / P[(size_t)j * r + j];
      __syncwarp();
      if (lane == 0) v[j] = yj;
      for (int i = lane; i < j; i += 32) v[i] = fma(-P[(size_t)j * r + i], yj, v[i]);
      __syncwarp();
    }
    }
    for (int j = lane; j < c; j += 32) Xc[perm ? perm[co + j] : co + j] = v[j];
  } else if (op == 1 /*OP_APPLY_QT*/ && full_q) {
    for (int i = lane; i < c; i += 32) Xc[co + i] = v[i];
    for (int i = lane; i < r - c; i += 32) Xc[n_cols + m1off + i] = v[c + i];
  } else {
    for (int i = lane; i < r; i += 32) Xc[ro + i] = v[i];
  }
}

}  // namespace qrk
