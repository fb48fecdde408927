// angular_inst.cu — instantiations of the block-angular kernels for ONE border width (-DQRK_M2=k).
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

template <int R, int C, bool PIV>
cudaError_t factor_t(const AngularArgs& a, cudaStream_t s) {
  using G = Cfg<R, C>;
  auto kernel = angular_factor_kernel<R, C, PIV, M2, TPB, G::U, G::NSTAGE, G::MINB>;
  const size_t smem = G::Smem::bytes;
  cudaError_t e = opt_in(kernel, smem);
  if (e != cudaSuccess) return e;
  kernel<<<a.grid, TPB, smem, s>>>(a.A_in, a.packed, a.tau, a.perm, a.J2, a.ldj, a.b, a.atop, a.y1, a.abot, a.partials, a.nb);
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
  *grid = sms * (per_sm > 0 ? per_sm : 1);
  return cudaSuccess;
}

template <int R, int C>
cudaError_t rhs_t(const AngularArgs& a, cudaStream_t s) {
  auto kernel = angular_rhs_kernel<R, C, M2, TPB, minb<R, C>()>;
  const size_t smem = ((size_t)TPB * (Group<R * C>::stride + Group<C>::stride) + (TPB / 32) * Tri<M2>::N) * 8;
  cudaError_t e = opt_in(kernel, smem);
  if (e != cudaSuccess) return e;
  kernel<<<a.grid, TPB, smem, s>>>(a.packed, a.tau, a.b, a.y1, a.abot, a.partials, a.nb);
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
    angular_root_kernel<M2, 512, true><<<1, 512, 0, s>>>(a.tris, a.tri_count, a.root_mode, a.out_tri, a.root, a.root_i, a.keep_rhs_only,
                                                         a.perm_tail, a.m1, xc);
  else
    angular_root_kernel<M2, 512, false><<<1, 512, 0, s>>>(a.tris, a.tri_count, a.root_mode, a.out_tri, a.root, a.root_i, a.keep_rhs_only,
                                                          a.perm_tail, a.m1, xc);
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
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, angular_rhs_kernel<R, C, M2, TPB, minb<R, C>()>);
  if (e == cudaSuccess) e = piv ? cudaFuncGetAttributes(&fa, angular_backsolve_kernel<R, C, M2, true, TPB>)
                                : cudaFuncGetAttributes(&fa, angular_backsolve_kernel<R, C, M2, false, TPB>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, angular_root_kernel<M2, 512, false>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, angular_root_kernel<M2, 512, true>);
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
