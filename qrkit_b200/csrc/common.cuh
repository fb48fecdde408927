// common.cuh — shared device helpers of the qrkit_b200 kernels (sm_100a only).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "qrkit_b200 is written for sm_100a (B200) only"
#endif

namespace qrk {

// ---------------------------------------------------------------------------------------------
// Shared-memory staging of "groups": a tile is `count` groups of N doubles, contiguous in global
// memory (the block-COO layout: consecutive blocks are adjacent).  In shared memory every group gets
// a stride of S doubles chosen so that thread t reading *its own* group t is bank-conflict free:
//   N even -> 16-byte accesses; a quarter-warp (8 threads x 16 B) is conflict free iff S/2 is odd;
//   N odd  ->  8-byte accesses; a half-warp (16 threads x 8 B) is conflict free iff S is odd.
// Global traffic is always fully coalesced 16-byte (8-byte for odd N) accesses.
// ---------------------------------------------------------------------------------------------
template <int N>
struct Group {
  static constexpr int vec = (N % 2 == 0) ? 2 : 1;
  static constexpr int stride = (N % 2 == 1) ? N : ((N % 4 == 2) ? N : N + 2);
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
// L2 eviction-priority hints (126 MB L2): streaming inputs are read with evict_first so that they do not displace the
// producer kernel's outputs, which are written with evict_last and are then still L2-resident when the consumer kernel of
// the same step reads them back (block-angular K1 -> K3).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cp_async16_hint(void* smem, const void* gmem, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async8_hint(void* smem, const void* gmem, uint64_t pol) {
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_global_hint(double* g, double v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(g), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_global_hint(double2* g, double2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(g), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ double2 ld_global_nc_hint(const double2* g, uint64_t pol) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(g), "l"(pol));
  return v;
}
__device__ __forceinline__ double2 ld_global_hint(const double2* g, uint64_t pol) {     // coherent path: the buffer may be written by this kernel
  double2 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(g), "l"(pol) : "memory");
  return v;
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int K>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(K) : "memory"); }

// global (contiguous) -> shared (strided groups), asynchronous (LDGSTS): no register staging.
// TPB threads copy a tile of up to TILE groups (`count` of them valid).
template <int N, int S, int TPB, int TILE = TPB>
__device__ __forceinline__ void stage_in_async(double* s, const double* __restrict__ g, int count) {
  constexpr int V = Group<N>::vec;
  if (count == TILE) {
    constexpr int total = TILE * N / V;
#pragma unroll
    for (int q0 = 0; q0 < total; q0 += TPB) {
      const int q = q0 + threadIdx.x;
      if (total % TPB == 0 || q < total) {
        const int d = q * V, grp = d / N, within = d - grp * N;
        if (V == 2) cp_async16(s + grp * S + within, g + d);
        else cp_async8(s + grp * S + within, g + d);
      }
    }
  } else {
    const int total = count * N / V + ((count * N) % V ? 1 : 0);
    for (int q = threadIdx.x; q < total; q += TPB) {
      const int d = q * V, grp = d / N, within = d - grp * N;
      if (V == 2) cp_async16(s + grp * S + within, g + d);   // count*N is even whenever N is
      else cp_async8(s + grp * S + within, g + d);
    }
  }
}

// as stage_in_async, every copy carrying an L2 cache policy
template <int N, int S, int TPB, int TILE = TPB>
__device__ __forceinline__ void stage_in_async_hint(double* s, const double* __restrict__ g, int count, uint64_t pol) {
  constexpr int V = Group<N>::vec;
  if (count == TILE) {
    constexpr int total = TILE * N / V;
#pragma unroll
    for (int q0 = 0; q0 < total; q0 += TPB) {
      const int q = q0 + threadIdx.x;
      if (total % TPB == 0 || q < total) {
        const int d = q * V, grp = d / N, within = d - grp * N;
        if (V == 2) cp_async16_hint(s + grp * S + within, g + d, pol);
        else cp_async8_hint(s + grp * S + within, g + d, pol);
      }
    }
  } else {
    const int total = count * N / V + ((count * N) % V ? 1 : 0);
    for (int q = threadIdx.x; q < total; q += TPB) {
      const int d = q * V, grp = d / N, within = d - grp * N;
      if (V == 2) cp_async16_hint(s + grp * S + within, g + d, pol);
      else cp_async8_hint(s + grp * S + within, g + d, pol);
    }
  }
}

// as stage_out, every store carrying an L2 cache policy
template <int N, int S, int TPB, int TILE = TPB>
__device__ __forceinline__ void stage_out_hint(double* __restrict__ g, const double* s, int count, uint64_t pol) {
  constexpr int V = Group<N>::vec;
  const int total = (count == TILE) ? TILE * N / V : count * N / V;
#pragma unroll 4
  for (int q = threadIdx.x; q < total; q += TPB) {
    const int d = q * V, grp = d / N, within = d - grp * N;
    if (V == 2) st_global_hint(reinterpret_cast<double2*>(g + d), *reinterpret_cast<const double2*>(s + grp * S + within), pol);
    else st_global_hint(g + d, s[grp * S + within], pol);
  }
}

// shared (strided groups) -> global (contiguous), coalesced vector stores.
template <int N, int S, int TPB, int TILE = TPB>
__device__ __forceinline__ void stage_out(double* __restrict__ g, const double* s, int count) {
  constexpr int V = Group<N>::vec;
  const int total = (count == TILE) ? TILE * N / V : count * N / V;
#pragma unroll 4
  for (int q = threadIdx.x; q < total; q += TPB) {
    const int d = q * V, grp = d / N, within = d - grp * N;
    if (V == 2) *reinterpret_cast<double2*>(g + d) = *reinterpret_cast<const double2*>(s + grp * S + within);
    else g[d] = s[grp * S + within];
  }
}

// int32 groups (column permutation): stride odd => conflict free 4-byte accesses.
template <int N>
struct GroupI32 { static constexpr int stride = (N % 2 == 1) ? N : N + 1; };

template <int N, int S, int TPB>
__device__ __forceinline__ void stage_out_i32(int* __restrict__ g, const int* s, int count) {
  const int total = count * N;
  for (int d = threadIdx.x; d < total; d += TPB) {
    const int grp = d / N, within = d - grp * N;
    g[d] = s[grp * S + within];
  }
}

// per-thread group <-> registers
template <int N>
__device__ __forceinline__ void load_group(double (&dst)[N], const double* s) {
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
      const double2 v = reinterpret_cast<const double2*>(s)[i];
      dst[2 * i] = v.x; dst[2 * i + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; i++) dst[i] = s[i];
  }
}
template <int N>
__device__ __forceinline__ void store_group(double* s, const double (&src)[N]) {
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; i++) reinterpret_cast<double2*>(s)[i] = make_double2(src[2 * i], src[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < N; i++) s[i] = src[i];
  }
}

// ---------------------------------------------------------------------------------------------
// FP64 reciprocal / reciprocal square root without the slow-path branches of the IEEE sequences:
// MUFU.RCP64H / MUFU.RSQ64H seed (20 good bits) + two Newton steps -> <= 1 ulp (not correctly rounded).
// The per-block QR needs, per reflector, norm, 1/beta and 1/(x0 - beta); with the IEEE sqrt and two
// IEEE divisions those three take ~40 instructions and three branches, here ~17 straight-line ones.
// Zero / infinite inputs keep the seed's inf / 0 (as the IEEE operations would give).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double x) {
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
  double e = fma(-x, r0, 1.0);      // seed: 19.9 good bits (measured, tools/seed_accuracy.cu) -> 40 -> 80
  double r = fma(r0, e, r0);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return (r0 == 0.0 || fabs(r0) == __longlong_as_double(0x7ff0000000000000LL)) ? r0 : r;
}
// returns 1/sqrt(x); sqrt_out = sqrt(x).  x must be >= 0.
__device__ __forceinline__ double fast_rsqrt(double x, double& sqrt_out) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double h = 0.5 * x;
  double y = y0;
#pragma unroll
  for (int it = 0; it < 2; it++) {     // seed: 20.0 good bits (measured) -> 39 -> 78
    const double t = y * y;
    const double e = fma(-h, t, 0.5);
    y = fma(y, e, y);
  }
  double s = x * y;
  s = fma(fma(-s, s, x), 0.5 * y, s);     // one Heron correction of the square root
  const bool special = (y0 == 0.0) || (y0 == __longlong_as_double(0x7ff0000000000000LL));
  sqrt_out = special ? ((x == 0.0) ? 0.0 : x) : s;   // x = 0 -> 0, x = inf -> inf
  return special ? y0 : y;
}

// Scalars of one Householder reflector (Eigen's makeHouseholder, [Eigen] Householder.h): x0 = pivot entry, tailSq =
// squared norm of the entries below.  beta = -sign(x0) sqrt(x0^2 + tailSq), inv = 1/(x0 - beta) (essential part = tail * inv),
// tau = (beta - x0)/beta; tailSq <= DBL_MIN (or `no_tail`) gives the identity: tau = 0, inv = 0, beta = x0.
// This is THE dependent chain of every column step (TSQR merges, banded chase, panel steps), so its depth is what is
// minimised (a dependent FP64 operation costs ~18 cycles on this part; r02 root trace: 750 cycles per merged column):
//   * the norm is taken from the rsqrt after ONE Newton step (20 -> 39 bits) and one Heron correction, which squares the error
//     again (-> 78 bits): the second Newton step of 1/norm is only needed by tau and runs beside the chain;
//   * the reciprocal's MUFU seed is taken from the 20-bit norm while those steps run, and is finished by ONE cubic step
//     r0 (1 + e + e^2), e = 1 - d r0 (|e| < 2^-19 -> 2^-57 relative) instead of two Newton steps.
// inv is ready 10 dependent operations after the rsqrt seed (14 before); all three results are correctly rounded to within
// one ulp of the two-step version (tests: R / tau against the oracle at 1e-12).
__device__ __forceinline__ void householder_scalars(double x0, double tailSq, bool no_tail, double& beta, double& inv, double& tau) {
  const double s = fma(x0, x0, tailSq);
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(s));
  const double ax = fabs(x0);
  double r0;
  {
    const double d0 = fma(s, y0, ax);            // |x0| + norm to ~20 bits
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d0));
  }
  const double h = 0.5 * s;
  double y1;
  { const double t = y0 * y0; const double e = fma(-h, t, 0.5); y1 = fma(y0, e, y0); }        // 39 bits
  double nrm = s * y1;
  nrm = fma(fma(-nrm, nrm, s), 0.5 * y1, nrm);   // Heron correction: full precision
  double y;                                       // 1/norm to full precision (tau only; off the chain)
  { const double t = y1 * y1; const double e = fma(-h, t, 0.5); y = fma(y1, e, y1); }
  const double d = ax + nrm;                     // |x0 - beta|
  const double e = fma(-d, r0, 1.0);
  const double r = fma(r0, fma(e, e, e), r0);    // r0 (1 + e + e^2)
  const bool neg = x0 < 0.0;
  const bool degenerate = no_tail || (tailSq <= DBL_MIN) || !(s < __longlong_as_double(0x7ff0000000000000LL));
  beta = degenerate ? x0 : (neg ? nrm : -nrm);
  inv = degenerate ? 0.0 : (neg ? -r : r);       // x0 - beta = sign(x0) (|x0| + norm), sign(0) = +
  tau = degenerate ? 0.0 : d * y;                // (beta - x0)/beta = (|x0| + norm)/norm
}

// splitmix64 counter-based generator shared with the test-suite (tests/helpers.py)
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ double synth_value(uint64_t seed, uint64_t block, uint64_t row, uint64_t col,
                                                       double lo, double hi) {
  const uint64_t u = splitmix64(seed ^ (block << 20) ^ (row << 10) ^ col);
  return lo + (hi - lo) * (double)(u >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace qrk
