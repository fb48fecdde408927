// angular_inst.cu — instantiations of the block-angular kernels for ONE border width (-DQRK_M2=k).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include "angular.cuh"
#include "angular_dispatch.hpp"

#ifndef QRK_M2
#error "compile with -DQRK_M2=<border columns>"
#endif

namespace qrk {
namespace {

constexpr int M2 = QRK_M2;
// tuning hooks for A/B builds (QRK_NVCC_EXTRA="-DQRK_ANG_TPB=.. -DQRK_ANG_U1=.. -DQRK_ANG_MINB=.."); the defaults are the measured best
#ifndef QRK_ANG_TPB
#define QRK_ANG_TPB 128
#endif
#ifndef QRK_ANG_U1
#define QRK_ANG_U1 2
#endif
constexpr int TPB = QRK_ANG_TPB;

// left block shapes available to the block-angular path (r > c)
#define QRK_ANGULAR_SHAPES(X) X(2, 1) X(3, 1) X(4, 2) X(7, 2)

// tile shape of the factor kernel: U blocks per thread so that one fold covers >= 2 residual rows where the
// footprint allows, and a double-buffered cp.async ring when two stages fit comfortably in shared memory
template <int R, int C>
struct Cfg {
  static constexpr int M1 = R - C;
  static constexpr int U = (M1 >= 2) ? 1 : QRK_ANG_U1;
  static constexpr size_t stage_bytes = (size_t)AngularSmem<R, C, M2, TPB, U, 1>::stage_doubles * 8;
  static constexpr int NSTAGE = (2 * stage_bytes <= 96 * 1024) ? 2 : 1;
  static constexpr int regs_doubles = R * C + U * M1 * (M2 + 1) + Tri<M2>::N;
#ifdef QRK_ANG_MINB
  static constexpr int MINB = QRK_ANG_MINB;
#else
  static constexpr int MINB = (regs_doubles <= 44) ? 3 : 2;
#endif
  using Smem = AngularSmem<R, C, M2, TPB, U, NSTAGE>;
};

template <int R, int C>
constexpr int minb() { return Cfg<R, C>::MINB; }

template <typename K>
cudaError_t opt_in(K kernel, size_t smem) {
  if (smem <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// 2 x 1 blocks: the direct (unstaged) kernel when every slice is a 16-byte aligned vector
// U points per fold and CTAs per SM: U = 4 at 168 registers (3 CTAs of 128 threads) is the measured best for the 5-column
// border of configs 1 / 3; wider borders carry a bigger triangle and more border words per point, so they take fewer points
// per fold to stay free of spills (ptxas -v: M2 = 6: U = 3, M2 = 7, 8: U = 2 with 2 CTAs)
#ifndef QRK_ANG_DIRECT_U
#define QRK_ANG_DIRECT_U (QRK_M2 <= 5 ? 4 : QRK_M2 == 6 ? 3 : 2)
#endif
#ifndef QRK_ANG_DIRECT_MINB
#define QRK_ANG_DIRECT_MINB (QRK_M2 <= 6 ? 3 : 2)
#endif
constexpr int kDirectU = QRK_ANG_DIRECT_U, kDirectMinB = QRK_ANG_DIRECT_MINB;
#ifndef QRK_ANG_K3_TPB
#define QRK_ANG_K3_TPB 256
#endif
#ifndef QRK_ANG_K3_PR
#define QRK_ANG_K3_PR 4
#endif
#ifndef QRK_ANG_K3_PS
#define QRK_ANG_K3_PS 4
#endif
// threads of the TSQR root: TPB * MergeFan::NT triangles are merged in two dependent warp merges (angular_root_kernel)
#ifdef QRK_ANG_ROOT_TPB
constexpr int kRootTpb = QRK_ANG_ROOT_TPB;
#else
constexpr int kRootTpb = 512;
#endif
constexpr int kK3Tpb = QRK_ANG_K3_TPB, kK3Pr = QRK_ANG_K3_PR, kK3Ps = QRK_ANG_K3_PS;
constexpr size_t kK3Smem = (size_t)kK3Ps * (M2 + 2) * kK3Tpb * sizeof(double);
inline bool direct_ok(const AngularArgs& a) {
  static const bool off = std::getenv("QRK_ANG_STAGED") != nullptr;     // A/B switch: force the staged kernel
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const long long span = (long long)(M2 + 1) * std::max<long long>(a.nb, a.ldj / 2) + 8ll * 148 * 16 * TPB;    // + a sweep of headroom for the loop counter
  return !off && a.r == 2 && a.c == 1 && al16(a.A_in) && al16(a.packed) && al16(a.J2) && al16(a.b) && (a.ldj % 2 == 0) && span < (1ll << 32);
}

template <bool PIV, bool ABOT>
cudaError_t factor_2x1(const AngularArgs& a, cudaStream_t s) {
  angular_factor_direct_kernel<PIV, M2, TPB, kDirectU, kDirectMinB, ABOT><<<a.grid, TPB, 0, s>>>(
      a.A_in, a.packed, a.tau, a.perm, a.J2, a.ldj, a.b, a.atop, a.y1, a.abot, a.partials, a.nb, a.grid);
  return cudaGetLastError();
}

template <int R, int C, bool PIV>
cudaError_t factor_t(const AngularArgs& a, cudaStream_t s) {
  using G = Cfg<R, C>;
  if constexpr (R == 2 && C == 1) {
    if (direct_ok(a)) return a.abot ? factor_2x1<PIV, true>(a, s) : factor_2x1<PIV, false>(a, s);
  }
  auto kernel = angular_factor_kernel<R, C, PIV, M2, TPB, G::U, G::NSTAGE, G::MINB>;
  const size_t smem = G::Smem::bytes;
  cudaError_t e = opt_in(kernel, smem);
  if (e != cudaSuccess) return e;
  int grid = a.grid;
  if constexpr (R == 2 && C == 1) {
    // a.grid was sized for the direct kernel's residency: the staged kernel runs one wave of ITS resident CTAs and the
    // partial triangles it does not produce are zero (a zero triangle is the identity of the TSQR merge)
    static int resident = 0;
    if (resident == 0) {
      int dev = 0, sms = 0, per_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TPB, smem);
      if (e != cudaSuccess) return e;
      resident = sms * (per_sm > 0 ? per_sm : 1);
    }
    const long long ntiles = (a.nb + G::Smem::TILE - 1) / G::Smem::TILE;
    grid = (int)std::max<long long>(1, std::min<long long>(std::min(grid, resident), ntiles));
    if (grid < a.grid) {
      e = cudaMemsetAsync(a.partials, 0, (size_t)a.grid * Tri<M2>::N * sizeof(double), s);     // (component-major: the unused columns are strided)
      if (e != cudaSuccess) return e;
    }
  }
  kernel<<<grid, TPB, smem, s>>>(a.A_in, a.packed, a.tau, a.perm, a.J2, a.ldj, a.b, a.atop, a.y1, a.abot, a.partials, a.nb, a.grid);
  return cudaGetLastError();
}

template <int R, int C, bool PIV>
cudaError_t max_grid_t(int* grid) {
  using G = Cfg<R, C>;
  auto kernel = angular_factor_kernel<R, C, PIV, M2, TPB, G::U, G::NSTAGE, G::MINB>;
  const size_t smem = G::Smem::bytes;
  cudaError_t e = opt_in(kernel, smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TPB, smem);
  if (e != cudaSuccess) return e;
  if constexpr (R == 2 && C == 1) {       // the direct kernel holds more CTAs per SM; the staged one copes with any grid
    int per_sm_direct = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_direct, angular_factor_direct_kernel<PIV, M2, TPB, kDirectU, kDirectMinB, false>, TPB, 0);
    if (e != cudaSuccess) return e;
    if (per_sm_direct > per_sm) per_sm = per_sm_direct;
  }
  *grid = sms * (per_sm > 0 ? per_sm : 1);
  return cudaSuccess;
}

template <int R, int C>
cudaError_t rhs_t(const AngularArgs& a, cudaStream_t s) {
  if constexpr (R == 2 && C == 1) {
    static const bool off = std::getenv("QRK_ANG_STAGED") != nullptr;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    if (!off && al16(a.packed) && al16(a.b) && (long long)(M2 + 2) * a.nb + 8ll * 148 * 16 * TPB < (1ll << 32)) {
      angular_rhs_direct_kernel<M2, TPB, kDirectU, kDirectMinB><<<a.grid, TPB, 0, s>>>(a.packed, a.tau, a.b, a.y1, a.abot, a.partials, a.nb, a.grid);
      return cudaGetLastError();
    }
  }
  auto kernel = angular_rhs_kernel<R, C, M2, TPB, minb<R, C>()>;
  const size_t smem = ((size_t)TPB * (Group<R * C>::stride + Group<C>::stride) + (TPB / 32) * Tri<M2>::N) * 8;
  cudaError_t e = opt_in(kernel, smem);
  if (e != cudaSuccess) return e;
  kernel<<<a.grid, TPB, smem, s>>>(a.packed, a.tau, a.b, a.y1, a.abot, a.partials, a.nb, a.grid);
  return cudaGetLastError();
}

template <int R, int C>
cudaError_t backsolve_t(const AngularArgs& a, cudaStream_t s) {
  // programmatic dependent launch behind the root kernel: see angular_backsolve_kernel
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((a.nb + TPB - 1) / TPB));
  cfg.blockDim = dim3(TPB);
  cfg.dynamicSmemBytes = (size_t)TPB * Group<R * C>::stride * 8;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = std::getenv("QRK_NO_PDL") == nullptr;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  const double* packed = a.packed; const int* perm = a.perm; const double* atop = a.atop; const double* y1 = a.y1; const double* root = a.root;
  double* x = a.x; long long nb = a.nb;
  if constexpr (R == 2 && C == 1) {        // (a 1-column block has the identity permutation: PERM is moot)
    static const bool old_k3 = std::getenv("QRK_ANG_K3_TILED") != nullptr;     // A/B switch: the generic one-point-per-thread kernel
    if (!old_k3) {
      auto k3 = angular_backsolve_direct_kernel<M2, kK3Tpb, kK3Pr, kK3Ps>;
      cudaError_t e = opt_in(k3, kK3Smem);          // (a per-device attribute: every launch, like the other kernels)
      if (e != cudaSuccess) return e;
      static int resident = 0;                      // one wave: the CTAs that are resident together (see the kernel)
      if (resident == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3, kK3Tpb, kK3Smem);
        if (e != cudaSuccess) return e;
        resident = sms * (per_sm > 0 ? per_sm : 1);
      }
      const long long batch = (long long)kK3Tpb * (kK3Pr + kK3Ps);
      cfg.gridDim = dim3((unsigned)std::max<long long>(1, std::min<long long>((a.nb + batch - 1) / batch, resident)));
      cfg.blockDim = dim3(kK3Tpb);
      cfg.dynamicSmemBytes = kK3Smem;
      return cudaLaunchKernelEx(&cfg, k3, packed, atop, y1, root, x, nb);
    }
  }
  if (a.piv) return cudaLaunchKernelEx(&cfg, angular_backsolve_kernel<R, C, M2, true, TPB>, packed, perm, atop, y1, root, x, nb);
  return cudaLaunchKernelEx(&cfg, angular_backsolve_kernel<R, C, M2, false, TPB>, packed, perm, atop, y1, root, x, nb);
}

bool shape_ok(int r, int c) {
#define X(R_, C_) if (r == R_ && c == C_) return true;
  QRK_ANGULAR_SHAPES(X)
#undef X
  return false;
}

int tile_blocks(int r, int c) {
  if (r == 2 && c == 1) return TPB;       // direct kernel: a CTA sweep covers TPB points
#define X(R_, C_) if (r == R_ && c == C_) return Cfg<R_, C_>::U * TPB;
  QRK_ANGULAR_SHAPES(X)
#undef X
  return TPB;
}

cudaError_t max_grid(int r, int c, bool piv, int* grid) {
#define X(R_, C_) if (r == R_ && c == C_) return piv ? max_grid_t<R_, C_, true>(grid) : max_grid_t<R_, C_, false>(grid);
  QRK_ANGULAR_SHAPES(X)
#undef X
  return cudaErrorInvalidValue;
}
cudaError_t factor(const AngularArgs& a, cudaStream_t s) {
#define X(R_, C_) if (a.r == R_ && a.c == C_) return a.piv ? factor_t<R_, C_, true>(a, s) : factor_t<R_, C_, false>(a, s);
  QRK_ANGULAR_SHAPES(X)
#undef X
  return cudaErrorInvalidValue;
}
cudaError_t rhs(const AngularArgs& a, cudaStream_t s) {
#define X(R_, C_) if (a.r == R_ && a.c == C_) return rhs_t<R_, C_>(a, s);
  QRK_ANGULAR_SHAPES(X)
#undef X
  return cudaErrorInvalidValue;
}
cudaError_t backsolve(const AngularArgs& a, cudaStream_t s) {
#define X(R_, C_) if (a.r == R_ && a.c == C_) return backsolve_t<R_, C_>(a, s);
  QRK_ANGULAR_SHAPES(X)
#undef X
  return cudaErrorInvalidValue;
}
cudaError_t root(const AngularArgs& a, cudaStream_t s) {
  AngularXchg xc;
  xc.peers = a.xchg_peers; xc.world = a.xchg_world; xc.rank = a.xchg_rank; xc.seq = a.xchg_seq; xc.err = a.xchg_err;
  if (a.xchg_timeout_ns) xc.timeout_ns = a.xchg_timeout_ns;
  if (a.root_mode == 2)
    angular_root_kernel<M2, kRootTpb, true><<<1, kRootTpb, 0, s>>>(a.tris, a.tri_count, a.tris_ld, a.root_mode, a.out_tri, a.root, a.root_i,
                                                                   a.keep_rhs_only, a.perm_tail, a.m1, xc);
  else
    angular_root_kernel<M2, kRootTpb, false><<<1, kRootTpb, 0, s>>>(a.tris, a.tri_count, a.tris_ld, a.root_mode, a.out_tri, a.root, a.root_i,
                                                                    a.keep_rhs_only, a.perm_tail, a.m1, xc);
  return cudaGetLastError();
}

// CUDA loads kernels lazily at their first launch, and loading can wait for running kernels: a root kernel that spins on a
// peer (mode 2) must never be the reason the peer's kernels cannot be loaded (two ranks emulated in one process would
// deadlock until the bounded spin gives up).  qrk_angular_p2p_attach therefore loads everything up front.
template <int R, int C>
cudaError_t preload_t(bool piv) {
  using G = Cfg<R, C>;
  cudaFuncAttributes fa;
  cudaError_t e = piv ? cudaFuncGetAttributes(&fa, angular_factor_kernel<R, C, true, M2, TPB, G::U, G::NSTAGE, G::MINB>)
                      : cudaFuncGetAttributes(&fa, angular_factor_kernel<R, C, false, M2, TPB, G::U, G::NSTAGE, G::MINB>);
  if constexpr (R == 2 && C == 1) {
    if (e == cudaSuccess) e = piv ? cudaFuncGetAttributes(&fa, angular_factor_direct_kernel<true, M2, TPB, kDirectU, kDirectMinB, false>)
                                  : cudaFuncGetAttributes(&fa, angular_factor_direct_kernel<false, M2, TPB, kDirectU, kDirectMinB, false>);
    if (e == cudaSuccess) e = piv ? cudaFuncGetAttributes(&fa, angular_factor_direct_kernel<true, M2, TPB, kDirectU, kDirectMinB, true>)
                                  : cudaFuncGetAttributes(&fa, angular_factor_direct_kernel<false, M2, TPB, kDirectU, kDirectMinB, true>);
  }
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, angular_rhs_kernel<R, C, M2, TPB, minb<R, C>()>);
  if constexpr (R == 2 && C == 1) {
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, angular_rhs_direct_kernel<M2, TPB, kDirectU, kDirectMinB>);
  }
  if (e == cudaSuccess) e = piv ? cudaFuncGetAttributes(&fa, angular_backsolve_kernel<R, C, M2, true, TPB>)
                                : cudaFuncGetAttributes(&fa, angular_backsolve_kernel<R, C, M2, false, TPB>);
  if constexpr (R == 2 && C == 1) {
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, angular_backsolve_direct_kernel<M2, kK3Tpb, kK3Pr, kK3Ps>);
  }
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, angular_root_kernel<M2, kRootTpb, false>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, angular_root_kernel<M2, kRootTpb, true>);
  return e;
}
cudaError_t preload(int r, int c, bool piv) {
#define X(R_, C_) if (r == R_ && c == C_) return preload_t<R_, C_>(piv);
  QRK_ANGULAR_SHAPES(X)
#undef X
  return cudaErrorInvalidValue;
}

const AngularVTable kTable = {M2, Tri<M2>::N, shape_ok, max_grid, tile_blocks, factor, rhs, root, backsolve, preload};

}  // namespace

#define QRK_CAT2(a, b) a##b
#define QRK_CAT(a, b) QRK_CAT2(a, b)
const AngularVTable* QRK_CAT(angular_vtable_m, QRK_M2)() { return &kTable; }

}  // namespace qrk
