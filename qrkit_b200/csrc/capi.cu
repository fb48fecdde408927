// capi.cu — the C ABI of qrkit_b200 (include/qrkit_b200.h): handle management, block-COO upload,
// kernel dispatch.  Host logic only; every floating-point operation of the hot path runs in the
// CUDA kernels of bd_small.cuh / bd_generic.cuh.  There is deliberately no CPU fallback.
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <tuple>
#include <new>
#include <numeric>

#include "bd_generic.cuh"
#include "bd_wy.cuh"
#include "dense_border.cuh"
#include "dense_blocked.cuh"
#include "dense_tri_reg.cuh"
#include "bd_small.cuh"
#include "export.cuh"
#include "solver.hpp"

using namespace qrk;

namespace qrk {
#define QRK_DECL_VT(k) const AngularVTable* angular_vtable_m##k();
QRK_DECL_VT(1) QRK_DECL_VT(2) QRK_DECL_VT(3) QRK_DECL_VT(4) QRK_DECL_VT(5) QRK_DECL_VT(6) QRK_DECL_VT(7) QRK_DECL_VT(8)
#undef QRK_DECL_VT
const AngularVTable* angular_vtable(int m2) {
  switch (m2) {
    case 1: return angular_vtable_m1();
    case 2: return angular_vtable_m2();
    case 3: return angular_vtable_m3();
    case 4: return angular_vtable_m4();
    case 5: return angular_vtable_m5();
    case 6: return angular_vtable_m6();
    case 7: return angular_vtable_m7();
    case 8: return angular_vtable_m8();
    default: return nullptr;
  }
}
}  // namespace qrk

namespace {

#define QRK_TRY_CUDA(h, expr)                                                                   \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      if (h) (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                    \
      return e__ == cudaErrorMemoryAllocation ? QRK_STATUS_ALLOC_FAILED : QRK_STATUS_CUDA_ERROR; \
    }                                                                                           \
  } while (0)

#define QRK_REQUIRE(h, cond, msg)                 \
  do {                                            \
    if (!(cond)) {                                \
      if (h) (h)->err = (msg);                    \
      return QRK_STATUS_INVALID_ARGUMENT;         \
    }                                             \
  } while (0)

struct DeviceGuard;
int peer_exchange_status(qrk_solver* h);   // QRK_STATUS_PEER_TIMEOUT once the fused exchange lost a peer (call after a stream sync)

struct DeviceGuard {
  int prev = -1;
  // Also clears a stale non-sticky error another library left in this thread (NCCL's lazy communicator set-up leaves
  // cudaErrorInvalidValue behind): every launch here is checked with cudaGetLastError and must only see its own.
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; (void)cudaGetLastError(); }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---- thread-per-block kernels: the (r, c) shapes instantiated at compile time -------------------
#define QRK_SMALL_SHAPES(X) X(2, 1) X(3, 1) X(4, 2) X(6, 3) X(7, 2) X(8, 2) X(8, 4) X(9, 2)
constexpr int kSmallTPB = 128;
constexpr int kSmallMinBlocks = 4;

// opt in to > 48 KB dynamic shared memory, once per (kernel instantiation, device)
// (`done` must be a static of the calling launch template: the kernel pointer TYPE is shared by all
// instantiations with the same signature)
template <typename K>
cudaError_t ensure_smem(K kernel, size_t bytes, bool (&done)[64]) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
  return e;
}

template <int R, int C, bool PIV, bool SOLVE>
cudaError_t launch_small_factor(const double* A, double* packed, double* tau, int* perm, const double* b, double* x,
                                long long nb, cudaStream_t s, long long block0) {
  auto kernel = bd_small_factor_kernel<R, C, PIV, SOLVE, kSmallTPB, kSmallMinBlocks>;
  constexpr size_t smem = SmallSmem<R, C, kSmallTPB>::bytes;
  static bool smem_opt_in[64] = {};
  cudaError_t attr = ensure_smem(kernel, smem, smem_opt_in);
  if (attr != cudaSuccess) return attr;
  const long long grid = (nb + kSmallTPB - 1) / kSmallTPB;
  kernel<<<(unsigned)grid, kSmallTPB, smem, s>>>(A, packed, tau, perm, b, x, nb, block0);
  return cudaGetLastError();
}

template <int R, int C, int OP, bool PERM>
cudaError_t launch_small_op_t(const double* packed, const double* tau, const int* perm, const double* B, long long ldb,
                              double* X, long long ldx, int nrhs, long long nb, long long n_cols, int full_q,
                              cudaStream_t s) {
  auto kernel = bd_small_op_kernel<R, C, OP, PERM, kSmallTPB, kSmallMinBlocks>;
  constexpr size_t smem = SmallOpSmem<R, C, kSmallTPB>::bytes;
  static bool smem_opt_in[64] = {};
  cudaError_t attr = ensure_smem(kernel, smem, smem_opt_in);
  if (attr != cudaSuccess) return attr;
  const long long grid = (nb + kSmallTPB - 1) / kSmallTPB;
  const int gy = (int)std::max<long long>(1, std::min<long long>(nrhs, (2 * 148 + grid - 1) / grid));   // fill the GPU when blocks are few
  kernel<<<dim3((unsigned)grid, (unsigned)gy), kSmallTPB, smem, s>>>(packed, tau, perm, B, ldb, X, ldx, nrhs, nb, n_cols, full_q);
  return cudaGetLastError();
}

bool small_shape_available(int r, int c) {
#define X(R_, C_) if (r == R_ && c == C_) return true;
  QRK_SMALL_SHAPES(X)
#undef X
  return false;
}

cudaError_t launch_small_factor_dyn(int r, int c, bool piv, bool solve, const double* A, double* packed, double* tau,
                                    int* perm, const double* b, double* x, long long nb, cudaStream_t s, long long block0 = 0) {
#define X(R_, C_)                                                                                     \
  if (r == R_ && c == C_) {                                                                           \
    if (piv) return solve ? launch_small_factor<R_, C_, true, true>(A, packed, tau, perm, b, x, nb, s, block0) \
                          : launch_small_factor<R_, C_, true, false>(A, packed, tau, perm, b, x, nb, s, block0); \
    return solve ? launch_small_factor<R_, C_, false, true>(A, packed, tau, perm, b, x, nb, s, block0)        \
                 : launch_small_factor<R_, C_, false, false>(A, packed, tau, perm, b, x, nb, s, block0);      \
  }
  QRK_SMALL_SHAPES(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t launch_small_op_dyn(int r, int c, int op, bool use_perm, const double* packed, const double* tau,
                                const int* perm, const double* B, long long ldb, double* X, long long ldx, int nrhs,
                                long long nb, long long n_cols, int full_q, cudaStream_t s) {
#define X(R_, C_)                                                                                                   \
  if (r == R_ && c == C_) {                                                                                         \
    if (op == OP_SOLVE)                                                                                             \
      return use_perm ? launch_small_op_t<R_, C_, OP_SOLVE, true>(packed, tau, perm, B, ldb, X, ldx, nrhs, nb, n_cols, full_q, s)  \
                      : launch_small_op_t<R_, C_, OP_SOLVE, false>(packed, tau, perm, B, ldb, X, ldx, nrhs, nb, n_cols, full_q, s); \
    if (op == OP_APPLY_QT)                                                                                          \
      return launch_small_op_t<R_, C_, OP_APPLY_QT, false>(packed, tau, perm, B, ldb, X, ldx, nrhs, nb, n_cols, full_q, s);        \
    return launch_small_op_t<R_, C_, OP_APPLY_Q, false>(packed, tau, perm, B, ldb, X, ldx, nrhs, nb, n_cols, full_q, s);           \
  }
  QRK_SMALL_SHAPES(X)
#undef X
  return cudaErrorInvalidValue;
}

// ---- generic (team-per-block) kernels ------------------------------------------------------------
template <int W>
cudaError_t launch_generic_factor_w(bool piv, bool solve, const BlockIndex& bi, const SizeClass& sc, const double* A,
                                    double* packed, double* tau, int* perm, const double* b, double* x, cudaStream_t s) {
#define LAUNCH(PIV, SOLVE)                                                                        \
  {                                                                                               \
    auto kernel = bd_generic_factor_kernel<W, PIV, SOLVE>;                                        \
    if (sc.smem > 48 * 1024) {                                                                    \
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc.smem); \
      if (e != cudaSuccess) return e;                                                             \
    }                                                                                             \
    kernel<<<(unsigned)sc.count, 32 * W, sc.smem, s>>>(bi, sc.d_ids, A, packed, tau, perm, b, x);  \
  }
  if (piv && solve) LAUNCH(true, true)
  else if (piv) LAUNCH(true, false)
  else if (solve) LAUNCH(false, true)
  else LAUNCH(false, false)
#undef LAUNCH
  return cudaGetLastError();
}

cudaError_t launch_generic_factor(bool piv, bool solve, const BlockIndex& bi, const SizeClass& sc, const double* A,
                                  double* packed, double* tau, int* perm, const double* b, double* x, cudaStream_t s) {
  switch (sc.warps) {
    case 1: return launch_generic_factor_w<1>(piv, solve, bi, sc, A, packed, tau, perm, b, x, s);
    case 4: return launch_generic_factor_w<4>(piv, solve, bi, sc, A, packed, tau, perm, b, x, s);
    default: return launch_generic_factor_w<8>(piv, solve, bi, sc, A, packed, tau, perm, b, x, s);
  }
}

constexpr size_t kMaxSmem = 227 * 1024;

inline bool ang(const qrk_solver* h) { return h->avt != nullptr || h->wide; }   // kind == QRK_BLOCK_ANGULAR

// ---- blocked compact-WY / DMMA kernel (bd_wy.cuh): unpivoted blocks wider than one panel -------------------------
template <int MR, int W>
cudaError_t launch_wy_factor_mw(bool solve, const BlockIndex& bi, const SizeClass& sc, const double* A, double* packed,
                                double* tau, const double* b, double* x, cudaStream_t s) {
#define LAUNCH(SOLVE)                                                                             \
  {                                                                                               \
    auto kernel = bd_wy_factor_kernel<MR, W, SOLVE>;                                              \
    if (sc.smem > 48 * 1024) {                                                                    \
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc.smem); \
      if (e != cudaSuccess) return e;                                                             \
    }                                                                                             \
    kernel<<<(unsigned)sc.count, 32 * W, sc.smem, s>>>(bi, sc.d_ids, A, packed, tau, b, x);        \
  }
  if (solve) LAUNCH(true)
  else LAUNCH(false)
#undef LAUNCH
  return cudaGetLastError();
}

cudaError_t launch_wy_factor(bool solve, const BlockIndex& bi, const SizeClass& sc, const double* A, double* packed,
                             double* tau, const double* b, double* x, cudaStream_t s) {
  const int key = sc.wy_mr * 10 + sc.warps;
  switch (key) {
    case 11: return launch_wy_factor_mw<1, 1>(solve, bi, sc, A, packed, tau, b, x, s);
    case 21: return launch_wy_factor_mw<2, 1>(solve, bi, sc, A, packed, tau, b, x, s);
    case 31: return launch_wy_factor_mw<3, 1>(solve, bi, sc, A, packed, tau, b, x, s);
    case 32: return launch_wy_factor_mw<3, 2>(solve, bi, sc, A, packed, tau, b, x, s);
    case 34: return launch_wy_factor_mw<3, 4>(solve, bi, sc, A, packed, tau, b, x, s);
    case 41: return launch_wy_factor_mw<4, 1>(solve, bi, sc, A, packed, tau, b, x, s);
    case 12: return launch_wy_factor_mw<1, 2>(solve, bi, sc, A, packed, tau, b, x, s);
    case 14: return launch_wy_factor_mw<1, 4>(solve, bi, sc, A, packed, tau, b, x, s);
    case 22: return launch_wy_factor_mw<2, 2>(solve, bi, sc, A, packed, tau, b, x, s);
    case 24: return launch_wy_factor_mw<2, 4>(solve, bi, sc, A, packed, tau, b, x, s);
    case 42: return launch_wy_factor_mw<4, 2>(solve, bi, sc, A, packed, tau, b, x, s);
    case 44: return launch_wy_factor_mw<4, 4>(solve, bi, sc, A, packed, tau, b, x, s);
    default: return cudaErrorInvalidValue;
  }
}

bool wy_eligible(int r, int c, bool piv) {
  return !piv && c > 8 && r >= c && wy_mr(r, c) != 0 && wy_smem_bytes(r, c) <= kMaxSmem;
}

int team_warps_for(int r, int c) {
  const long long e = (long long)r * c;
  if (e <= 1024) return 1;
  if (e <= 4096) return 4;
  return 8;
}

BlockIndex block_index(const qrk_solver* h) {
  BlockIndex bi;
  bi.rows = h->uniform ? nullptr : h->d_rows;
  bi.cols = h->d_cols; bi.voff = h->d_voff; bi.roff = h->d_roff; bi.coff = h->d_coff;
  bi.ur = h->ur; bi.uc = h->uc;
  return bi;
}

void free_dev(qrk_solver* h) {
  for (auto* g : {&h->sg_wide, &h->sg_tsqr, &h->sg_solve}) { if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; } g->seen = 0; }
  auto F = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
  F(h->d_rows); F(h->d_cols); F(h->d_voff); F(h->d_roff); F(h->d_coff);
  if (h->own_values) F(h->d_values);
  h->d_values = nullptr;
  F(h->d_tau); F(h->d_perm); F(h->d_b); F(h->d_x);
  F(h->d_rband); F(h->d_btau); F(h->d_ythin); F(h->d_gband); F(h->d_gy); F(h->d_cvec); F(h->d_ctau);
  F(h->d_wx); F(h->d_wupd); F(h->d_wdir); F(h->d_wtau2); F(h->d_wscal); F(h->d_wtau1); F(h->d_wT); F(h->d_wtri); F(h->d_wpart); F(h->d_xchg); F(h->d_xchg_peers); F(h->d_xchg_err); F(h->d_xchg_seq); F(h->d_wperm); F(h->d_wiscal);
  F(h->d_gwin); F(h->d_gcol0); F(h->d_gpacked); F(h->d_gtau); F(h->d_gcomp);
  F(h->d_q2); F(h->d_q2tau); F(h->d_q2sign); F(h->d_q2scr); F(h->d_qtmp); F(h->d_qthin); F(h->d_q2iscr);
  F(h->d_border_own); F(h->d_atop); F(h->d_y1); F(h->d_abot); F(h->d_partials); F(h->d_tri); F(h->d_root); F(h->d_root_i);
  for (auto& sc : h->classes) if (sc.d_ids) cudaFree(sc.d_ids);
  h->classes.clear();
  for (cudaEvent_t e : h->pipe_events) cudaEventDestroy(e);
  h->pipe_events.clear();
  if (h->s_in) { cudaStreamDestroy(h->s_in); h->s_in = nullptr; }
  for (cudaEvent_t e : h->aux_events) cudaEventDestroy(e);
  h->aux_events.clear();
  if (h->s_aux) { cudaStreamDestroy(h->s_aux); h->s_aux = nullptr; }
  if (h->s_out) { cudaStreamDestroy(h->s_out); h->s_out = nullptr; }
}

int ensure_buffer(qrk_solver* h, double*& p, size_t& cap, size_t n) {
  if (cap >= n && p) return QRK_STATUS_OK;
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
  QRK_TRY_CUDA(h, cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(double)));
  cap = n;
  return QRK_STATUS_OK;
}

int ensure_own_values(qrk_solver* h) {
  if (h->own_values && h->d_values) return QRK_STATUS_OK;
  h->d_values = nullptr;
  QRK_TRY_CUDA(h, cudaMalloc(&h->d_values, std::max<long long>(h->total_values, 1) * sizeof(double)));
  h->own_values = true;
  return QRK_STATUS_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// One factorisation pass over all blocks (optionally fused with the solve of one right-hand side).
// A_in: where the blocks are read from (h->d_values itself: in place; a caller's device buffer: out of
// place, no staging copy); the packed factors always land in h->d_values.
int run_factor(qrk_solver* h, const double* A_in, const double* d_b, double* d_x) {
  const bool piv = h->desc.pivoting == QRK_PIVOT_COLPIV;
  const bool solve = d_b != nullptr;
  if (h->nb == 0) return QRK_STATUS_OK;
  if (h->small_path) {
    QRK_TRY_CUDA(h, launch_small_factor_dyn(h->ur, h->uc, piv, solve, A_in, h->d_values, h->d_tau, h->d_perm, d_b,
                                            d_x, h->nb, h->stream));
    h->launches++;
  } else {
    const BlockIndex bi = block_index(h);
    for (const auto& sc : h->classes) {
      if (sc.wy_mr) {
        QRK_TRY_CUDA(h, launch_wy_factor(solve, bi, sc, A_in, h->d_values, h->d_tau, d_b, d_x, h->stream));
      } else {
        QRK_TRY_CUDA(h, launch_generic_factor(piv, solve, bi, sc, A_in, h->d_values, h->d_tau, h->d_perm, d_b, d_x,
                                              h->stream));
      }
      h->launches++;
    }
  }
  return QRK_STATUS_OK;
}

int run_op(qrk_solver* h, int op, const double* d_B, long long ldb, double* d_X, long long ldx, int nrhs) {
  if (h->nb == 0 || nrhs == 0) return QRK_STATUS_OK;
  const bool piv = h->desc.pivoting == QRK_PIVOT_COLPIV;
  const int full_q = h->desc.q_format == QRK_FULL_Q ? 1 : 0;
  const bool in_full_layout = (op == OP_APPLY_Q) && full_q;
  const long long q_nstart = ang(h) ? h->sum_cols : h->n_cols;   // N_start = mat.cols() of the block-diagonal matrix (:430)
  const bool vec_ok = aligned16(d_B) && aligned16(d_X) && (ldb % 2 == 0) && (ldx % 2 == 0) &&
                      (!(in_full_layout || (op == OP_APPLY_QT && full_q)) || (q_nstart % 2 == 0));
  if (h->small_path && vec_ok) {
    QRK_TRY_CUDA(h, launch_small_op_dyn(h->ur, h->uc, op, piv, h->d_values, h->d_tau, h->d_perm, d_B, ldb, d_X, ldx, nrhs,
                                        h->nb, q_nstart, full_q, h->stream));
  } else {
    constexpr int WPC = 4;
    const long long items = h->nb * (long long)nrhs;
    const long long grid = (items + WPC - 1) / WPC;
    const size_t smem = (size_t)WPC * h->max_r * sizeof(double);
    auto kernel = bd_generic_op_kernel<WPC>;
    if (smem > 48 * 1024) QRK_TRY_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<(unsigned)grid, 32 * WPC, smem, h->stream>>>(block_index(h), h->nb, h->d_values, h->d_tau,
                                                          piv ? h->d_perm : nullptr, d_B, ldb, d_X, ldx, nrhs, q_nstart, op,
                                                          full_q, h->max_r);
    QRK_TRY_CUDA(h, cudaGetLastError());
  }
  h->launches++;
  return QRK_STATUS_OK;
}

// rows of Q / R not covered by a block (BlockDiagonalSparseQR.h:530-533): Q(i,i) = 1, so Q^T b and
// Q b copy b there.  In the FullQ layout these rows sit at the same index i (columns n_cols+m1 .. map
// to themselves only when the index rule says so; the reference writes Q(i,i)=1 literally).
__global__ void copy_tail_kernel(const double* B, long long ldb, double* Y, long long ldy, int nrhs, long long from, long long to) {
  const long long n = to - from;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * nrhs; i += (long long)gridDim.x * blockDim.x) {
    const long long col = i / n, row = from + i % n;
    Y[col * ldy + row] = B[col * ldb + row];
  }
}

__global__ void iota_kernel(int* p, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = (int)i;
}

// ---- ellipse-fitting functor on the device (examples/ellipse_fitting.cpp:85-113; bench/bench_sparse_qr_extra.cpp:68-114)
__global__ void ellipse_points_kernel(double* __restrict__ px, double* __restrict__ py, long long n, double a, double b, double x0,
                                      double y0, double r) {
  const double incr = 1.3 * 3.14159265358979323846 / (double)n, cr = cos(r), sr = sin(r);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double st, ct;
    sincos((double)i * incr, &st, &ct);
    px[i] = x0 + a * ct * cr - b * st * sr;
    py[i] = y0 + a * ct * sr + b * st * cr;
  }
}

// one thread per point: 2x1 block of J1, its two rows of the 5 border columns, -f, and the cost
__global__ void __launch_bounds__(256) ellipse_assemble_kernel(const double* __restrict__ px, const double* __restrict__ py,
                                                              const double* __restrict__ params, long long n, double* __restrict__ J1,
                                                              double* __restrict__ J2, double* __restrict__ rhs, double* cost) {
  const double a = params[n], b = params[n + 1], x0 = params[n + 2], y0 = params[n + 3], r = params[n + 4];
  double sr, cr;
  sincos(r, &sr, &cr);
  double local = 0.0;
  const long long ld = 2 * n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double st, ct;
    sincos(params[i], &st, &ct);
    const double fx = px[i] - (a * ct * cr - b * st * sr + x0);
    const double fy = py[i] - (a * ct * sr + b * st * cr + y0);
    reinterpret_cast<double2*>(J1)[i] = make_double2(a * cr * st + b * sr * ct, a * sr * st - b * cr * ct);
    reinterpret_cast<double2*>(J2)[i] = make_double2(-ct * cr, -ct * sr);
    reinterpret_cast<double2*>(J2 + ld)[i] = make_double2(st * sr, -st * cr);
    reinterpret_cast<double2*>(J2 + 2 * ld)[i] = make_double2(-1.0, 0.0);
    reinterpret_cast<double2*>(J2 + 3 * ld)[i] = make_double2(0.0, -1.0);
    reinterpret_cast<double2*>(J2 + 4 * ld)[i] = make_double2(a * ct * sr + b * st * cr, -a * ct * cr + b * st * sr);
    reinterpret_cast<double2*>(rhs)[i] = make_double2(-fx, -fy);
    local = fma(fx, fx, fma(fy, fy, local));
  }
  if (cost) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ double sred[8];
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < 8; w++) s += sred[w];
      atomicAdd(cost, s);
    }
  }
}

__global__ void synth_fill_kernel(double* out, uint64_t seed, long long block0, long long nb, int r, int c, double lo, double hi) {
  const long long per = (long long)r * (c > 0 ? c : 1);
  const long long total = nb * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long blk = i / per;
    const int w = (int)(i - blk * per);
    if (c > 0) { const int col = w / r, row = w - col * r; out[i] = synth_value(seed, (uint64_t)(block0 + blk), row, col, lo, hi); }
    else out[i] = synth_value(seed, (uint64_t)(block0 * r + i), 0, 0, lo, hi);
  }
}



// ---- block angular ---------------------------------------------------------------------------------
// border columns of R = [R1, Atop P2; 0, R2] (makeR, BlockAngularSparseQR.h:296-305)
__global__ void export_angular_border_kernel(const double* __restrict__ atop, long long ld_atop, const double* __restrict__ root,
                                             const int* __restrict__ root_i, long long m1, int m2, long long base,
                                             int* __restrict__ outer, int* __restrict__ inner, double* __restrict__ vals) {
  const long long per_col_max = m1 + m2;
  const long long total = (long long)m2 * per_col_max;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / per_col_max);
    const long long r = i - (long long)c * per_col_max;
    const long long start = base + (long long)c * m1 + (long long)c * (c + 1) / 2;
    if (r == 0) outer[m1 + c] = (int)start;
    if (r < m1) {
      inner[start + r] = (int)r;
      vals[start + r] = atop[(long long)root_i[c] * ld_atop + r];
    } else if (r - m1 <= c) {
      inner[start + r] = (int)r;
      vals[start + r] = root[(long long)c * m2 + (r - m1)];
    }
    if (i == 0) outer[m1 + m2] = (int)(base + (long long)m2 * m1 + (long long)m2 * (m2 + 1) / 2);
  }
}

AngularArgs angular_args(qrk_solver* h) {
  AngularArgs a;
  a.r = h->ur; a.c = h->uc; a.piv = h->desc.pivoting == QRK_PIVOT_COLPIV; a.nb = h->nb;
  a.packed = h->d_values; a.tau = h->d_tau; a.perm = h->d_perm;
  a.J2 = h->d_border; a.ldj = h->ld_border;
  a.atop = h->d_atop; a.y1 = h->d_y1; a.abot = nullptr;
  a.partials = h->d_partials;
  const long long tile = h->avt->tile_blocks(h->ur, h->uc);
  const long long ntiles = (h->nb + tile - 1) / tile;
  a.grid = (int)std::max<long long>(1, std::min<long long>(h->a_grid, ntiles));
  a.tris = h->d_partials; a.tri_count = a.grid; a.tris_ld = a.grid;
  a.out_tri = h->d_tri; a.root = h->d_root; a.root_i = h->d_root_i;
  a.perm_tail = h->d_perm + h->sum_cols; a.m1 = (int)h->sum_cols;
  return a;
}

// TSQR root (+ back substitution when a right-hand side is present).  world > 1: either the triangles are exchanged with the
// peers inside the root kernel (after qrk_angular_p2p_attach), or the call stops at the per-GPU triangle and
// qrk_angular_merge finishes after the caller's all-gather.
int angular_root_and_back(qrk_solver* h, AngularArgs& a, bool have_rhs, double* d_x, int keep_rhs_only) {
  if (h->world > 1 && h->xchg_rank < 0) {
    a.root_mode = 0;
    QRK_TRY_CUDA(h, h->avt->root(a, h->stream));
    h->launches++;
    h->root_done = false;
    h->pending = true;
    h->pending_keep_rhs_only = keep_rhs_only;
    h->pending_x = have_rhs ? d_x : nullptr;
    h->pending_space = QRK_DEVICE;     // d_x is a device pointer here; the host-buffer entry points override both
    return QRK_STATUS_OK;
  }
  a.root_mode = 1;
  if (h->world > 1) {
    a.root_mode = 2;
    a.xchg_peers = h->d_xchg_peers; a.xchg_world = h->world; a.xchg_rank = h->xchg_rank; a.xchg_err = h->d_xchg_err;
    a.xchg_seq = h->d_xchg_seq; a.xchg_timeout_ns = h->xchg_timeout_ns;
  }
  a.keep_rhs_only = keep_rhs_only;
  QRK_TRY_CUDA(h, h->avt->root(a, h->stream));
  h->launches++;
  h->root_done = true;
  if (have_rhs) {
    a.x = d_x;
    QRK_TRY_CUDA(h, h->avt->backsolve(a, h->stream));
    h->launches++;
  }
  return QRK_STATUS_OK;
}

// ---- block angular, dense right block in global memory (dense_border.cuh) ------------------------------------------
int run_op(qrk_solver* h, int op, const double* d_B, long long ldb, double* d_X, long long ldx, int nrhs);
BandedArgs banded_args(qrk_solver* h);
int banded_run(qrk_solver* h, const double* A_in, const double* d_b, double* d_x);

DenseBorder wide_desc(qrk_solver* h, int nrhs) {
  DenseBorder d;
  d.A = h->d_wx + h->sum_cols;            // the complement rows of Q1^T [J2 | b] (rows [m1, n) for a block-diagonal left block)
  d.ld = h->w_ld; d.N = h->w_N; d.Nrule = h->n_rows - h->sum_cols; d.M = h->m2; d.nrhs = nrhs;
  d.pivot = h->desc.right_solver == QRK_RIGHT_UNPIVOTED ? 0 : h->desc.right_solver == QRK_RIGHT_THIN_SPARSE ? 2 : 1;
  d.upd = h->d_wupd; d.dir = h->d_wdir; d.tau = h->d_wtau2; d.perm = h->d_wperm; d.scal = h->d_wscal; d.iscal = h->d_wiscal;
  return d;
}

// The M x (M + nrhs) triangle left by the blocked first stage, as the matrix of the ColPiv second stage.
DenseBorder wide_tri_desc(qrk_solver* h, int nrhs) {
  DenseBorder d = wide_desc(h, nrhs);
  d.Nrule = d.N;                          // Eigen's threshold / rank rules count the rows of the tall residual
  d.A = h->d_wtri; d.ld = h->m2; d.N = h->m2;
  return d;
}

// rank / root record / P_c tail, then (with a right-hand side) y2, x2, x1 = P1 R1^-1 (ytop - Atop x2)
int wide_back(qrk_solver* h, bool have_rhs, double* d_x) {
  const bool two_stage = h->wide_blocked && h->desc.right_solver != QRK_RIGHT_UNPIVOTED;
  const DenseBorder d = two_stage ? wide_tri_desc(h, have_rhs ? 1 : 0) : wide_desc(h, have_rhs ? 1 : 0);
  const long long n = h->w_ld, m1 = h->sum_cols;
  const int M = h->m2;
  double* rhs_col = h->d_wx + (long long)M * n;
  const double* z = two_stage ? h->d_wtri + (size_t)M * M : rhs_col + m1;
  dense_finish_kernel<1024><<<1, 1024, (size_t)M * sizeof(double), h->stream>>>(d, have_rhs ? z : nullptr, h->d_root, h->d_root_i,
                                                                        h->d_perm + m1, (int)m1, have_rhs ? d_x + m1 : nullptr);
  QRK_TRY_CUDA(h, cudaGetLastError());
  h->launches++;
  h->root_done = true;
  if (have_rhs && m1 > 0) {
    dense_top_kernel<<<(unsigned)((m1 + 255) / 256), 256, (size_t)M * sizeof(double), h->stream>>>(h->d_wx, n, m1, M, h->d_root + (size_t)M * M + 2 * M,
                                                                                         rhs_col);
    QRK_TRY_CUDA(h, cudaGetLastError());
    if (h->left_banded) {                 // x1 = R1^-1 ytop with the band R1 (BandedBlockedSparseQR.h:299-304)
      BandedArgs a = banded_args(h);
      a.y = rhs_col; a.x = d_x;
      QRK_TRY_CUDA(h, h->bvt->backsolve(a, h->stream));
    } else {
      const bool piv = h->desc.pivoting == QRK_PIVOT_COLPIV;
      bd_rsolve_kernel<4><<<(unsigned)((h->nb + 3) / 4), 128, (size_t)4 * h->max_c * sizeof(double), h->stream>>>(
          block_index(h), h->nb, h->d_values, piv ? h->d_perm : nullptr, rhs_col, d_x, h->max_c);
      QRK_TRY_CUDA(h, cudaGetLastError());
    }
    h->launches += 2;
  }
  return QRK_STATUS_OK;
}

// Eigen's ColPivHouseholderQR (or, pivot = 0, HouseholderQR) column by column on d (dense_border.cuh)
int wide_unblocked(qrk_solver* h, const DenseBorder& d) {
  const int M = d.M;
  dense_norms_kernel<<<M, 256, 0, h->stream>>>(d);
  dense_prep_kernel<<<1, 256, 0, h->stream>>>(d);
  h->launches += 2;
  const int size = (int)std::min<long long>(d.N, M);
  for (int k = 0; k < size; k++) {                                     // rightSolver.compute(J2.bottomRows(...)) (:368)
    dense_piv_kernel<1024><<<1, 1024, 0, h->stream>>>(d, k);
    const int ncols = M - k - 1 + d.nrhs;
    if (ncols > 0) dense_upd_kernel<128><<<ncols, 128, 0, h->stream>>>(d, k);
    h->launches += ncols > 0 ? 2 : 1;
  }
  QRK_TRY_CUDA(h, cudaGetLastError());
  return QRK_STATUS_OK;
}


// BlockedThinSparseQR as the right solver (BlockedThinSparseQR.h:105-166): panels of `s` column positions; inside a panel
// Eigen's ColPivHouseholderQR column by column (dense_piv_kernel restricted to the panel, dense_upd_kernel on EVERY column to
// the right — updateMat(idxCol, cols), :264 — the right-hand side and the deferred columns included).  The panel's nonzero-pivot
// count comes back to the host once per panel; columns with a zero pivot are rotated behind the unprocessed ones (:250-255),
// so that "position = diagonal row" keeps holding for the next panel, which starts at row m_nonzeroPivots (:234).
int wide_thin_sparse(qrk_solver* h, DenseBorder d) {
  const int M = d.M;
  const int s = h->desc.reserved[1] > 0 ? h->desc.reserved[1] : 2;
  thin_init_kernel<<<1, 256, 0, h->stream>>>(d);
  h->launches++;
  h->thin_deferred = false;
  int nzp = 0, u1 = M;                      // [0, nzp) done, [nzp, u1) unprocessed in their original order, [u1, M) deferred
  double* tmp = nullptr;
  while (nzp < u1 && nzp < d.N) {
    const int k0 = nzp, left = u1 - k0;
    const int pc = (left <= s) ? left : s;  // updateBlockInfo (:198-236): the last panel takes what is left
    d.pend = k0 + pc;
    thin_panel_prep_kernel<<<1, 256, 0, h->stream>>>(d, k0, pc);
    h->launches++;
    const int steps = (int)std::min<long long>(pc, d.N - k0);
    for (int k = k0; k < k0 + steps; k++) {
      dense_piv_kernel<1024><<<1, 1024, 0, h->stream>>>(d, k);
      const int ncols = M - k - 1 + d.nrhs;
      if (ncols > 0) dense_upd_kernel<128><<<ncols, 128, 0, h->stream>>>(d, k);
      h->launches += ncols > 0 ? 2 : 1;
    }
    int nz_abs = k0 + pc;
    QRK_TRY_CUDA(h, cudaMemcpyAsync(&nz_abs, d.iscal + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
    if (steps < pc && nz_abs > k0 + steps) nz_abs = k0 + steps;      // fewer rows than panel columns: the rest has no pivot
    const int z = k0 + pc - nz_abs;
    if (z > 0) {                            // defer: zero below their own step, then [nz_abs, nz_abs + z) -> [u1 - z, u1)
      h->thin_deferred = true;
      for (int c = nz_abs; c < k0 + pc; c++) thin_clear_below_kernel<<<32, 256, 0, h->stream>>>(d, c, std::min(c, k0 + steps - 1));
      // [nz_abs, nz_abs + z) go to the very end, everything behind them (the unprocessed columns and the columns deferred
      // earlier, which stay in deferral order: m_zeroColPermIdxs, :253-255) shifts left by z
      const int moved = M - (nz_abs + z);
      if (moved > 0) {
        if (!tmp) QRK_TRY_CUDA(h, cudaMalloc(&tmp, (size_t)s * d.N * sizeof(double)));
        QRK_TRY_CUDA(h, cudaMemcpy2DAsync(tmp, d.N * sizeof(double), d.A + (long long)nz_abs * d.ld, d.ld * sizeof(double), d.N * sizeof(double), z,
                                          cudaMemcpyDeviceToDevice, h->stream));
        for (int c = nz_abs + z; c < M; c++)
          QRK_TRY_CUDA(h, cudaMemcpyAsync(d.A + (long long)(c - z) * d.ld, d.A + (long long)c * d.ld, d.N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        QRK_TRY_CUDA(h, cudaMemcpy2DAsync(d.A + (long long)(M - z) * d.ld, d.ld * sizeof(double), tmp, d.N * sizeof(double), d.N * sizeof(double), z,
                                          cudaMemcpyDeviceToDevice, h->stream));
        std::vector<int> perm(M);
        QRK_TRY_CUDA(h, cudaMemcpyAsync(perm.data(), d.perm, M * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
        std::rotate(perm.begin() + nz_abs, perm.begin() + nz_abs + z, perm.end());
        QRK_TRY_CUDA(h, cudaMemcpyAsync(d.perm, perm.data(), M * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
      }
      u1 -= z;
    }
    nzp = nz_abs;
  }
  QRK_TRY_CUDA(h, cudaMemcpyAsync(d.iscal, &nzp, sizeof(int), cudaMemcpyHostToDevice, h->stream));   // rank() = m_nonzeroPivots (:281)
  QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
  if (tmp) cudaFree(tmp);
  QRK_TRY_CUDA(h, cudaGetLastError());
  return QRK_STATUS_OK;
}

// Blocked compact-WY QR of the tall residual (dense_blocked.cuh): one cluster launch per 8-column panel + DMMA update
// launches over the trailing columns (the right-hand side included).  Look-ahead over two streams: the update of the NEXT
// panel's 8 columns and that panel's factorisation run on the auxiliary stream beside the update of the remaining columns.
constexpr int kWideRowSplit = 8;   // row ranges per column block of the trailing update

cudaError_t launch_dense_panel(const DenseBlocked& b, int k0, int pw, cudaStream_t s) {
  const long long rows = b.N - k0;
  const int rpt = (int)((rows + kDbCluster * kDbThreads - 1) / (kDbCluster * kDbThreads));
  static bool o1[64] = {}, o2[64] = {}, o4[64] = {};
  cudaError_t e;
  if (rpt <= 1) {
    if ((e = ensure_smem(dense_panel_kernel<1>, kDbPanelSmem, o1)) != cudaSuccess) return e;
    dense_panel_kernel<1><<<kDbCluster, kDbThreads, kDbPanelSmem, s>>>(b, k0, pw);
  } else if (rpt <= 2) {
    if ((e = ensure_smem(dense_panel_kernel<2>, kDbPanelSmem, o2)) != cudaSuccess) return e;
    dense_panel_kernel<2><<<kDbCluster, kDbThreads, kDbPanelSmem, s>>>(b, k0, pw);
  } else {
    if ((e = ensure_smem(dense_panel_kernel<4>, kDbPanelSmem, o4)) != cudaSuccess) return e;
    dense_panel_kernel<4><<<kDbCluster, kDbThreads, kDbPanelSmem, s>>>(b, k0, pw);
  }
  return cudaGetLastError();
}

int wide_blocked_qr(qrk_solver* h, const DenseBorder& d) {
  DenseBlocked b;
  b.A = d.A; b.ld = d.ld; b.N = d.N; b.M = d.M; b.ncols = d.M + d.nrhs;
  b.tau = d.pivot ? h->d_wtau1 : d.tau; b.T = h->d_wT;
  const int P = (d.M + 7) / 8;
  if (!h->s_aux) QRK_TRY_CUDA(h, cudaStreamCreateWithFlags(&h->s_aux, cudaStreamNonBlocking));
  while ((int)h->aux_events.size() < 2 * P + 2) {
    cudaEvent_t e;
    QRK_TRY_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->aux_events.push_back(e);
  }
  cudaStream_t main = h->stream, aux = h->s_aux;
  auto evP = [&](int p) { return h->aux_events[2 * p]; };       // panel p factored (recorded on aux)
  auto evU = [&](int p) { return h->aux_events[2 * p + 1]; };   // update p of the far columns done (recorded on main)
  cudaEvent_t fork = h->aux_events[2 * P], join = h->aux_events[2 * P + 1];
  QRK_TRY_CUDA(h, cudaEventRecord(fork, main));
  QRK_TRY_CUDA(h, cudaStreamWaitEvent(aux, fork, 0));
  if (launch_dense_panel(b, 0, std::min(8, d.M), aux) != cudaSuccess) return QRK_STATUS_UNSUPPORTED;   // no room for the cluster: nothing was modified yet
  QRK_TRY_CUDA(h, cudaEventRecord(evP(0), aux));
  h->launches++;
  for (int p = 0; p < P; p++) {
    const int k0 = 8 * p, pw = std::min(8, d.M - k0);
    const int ntrail = b.ncols - (k0 + pw);
    const int nblocks = (ntrail + 7) / 8;
    QRK_TRY_CUDA(h, cudaStreamWaitEvent(main, evP(p), 0));
    if (nblocks > 0) {
      if (p > 0) QRK_TRY_CUDA(h, cudaStreamWaitEvent(aux, evU(p - 1), 0));
      dense_wy_w_kernel<<<dim3(1, kWideRowSplit), 32 * kWyWarps, 0, aux>>>(b, k0, pw, 0, h->d_wpart);        // the next panel's columns first
      dense_wy_apply_kernel<<<dim3(1, kWideRowSplit), 32 * kWyWarps, 0, aux>>>(b, k0, pw, 0, h->d_wpart);
      h->launches += 2;
      if (p + 1 < P) {
        QRK_TRY_CUDA(h, launch_dense_panel(b, k0 + 8, std::min(8, d.M - k0 - 8), aux));
        QRK_TRY_CUDA(h, cudaEventRecord(evP(p + 1), aux));
        h->launches++;
      }
      if (nblocks > 1) {
        dense_wy_w_kernel<<<dim3(nblocks - 1, kWideRowSplit), 32 * kWyWarps, 0, main>>>(b, k0, pw, 1, h->d_wpart);
        dense_wy_apply_kernel<<<dim3(nblocks - 1, kWideRowSplit), 32 * kWyWarps, 0, main>>>(b, k0, pw, 1, h->d_wpart);
        h->launches += 2;
      }
      QRK_TRY_CUDA(h, cudaEventRecord(evU(p), main));
    }
  }
  QRK_TRY_CUDA(h, cudaEventRecord(join, aux));
  QRK_TRY_CUDA(h, cudaStreamWaitEvent(main, join, 0));
  QRK_TRY_CUDA(h, cudaGetLastError());
  return QRK_STATUS_OK;
}

// the blocked first stage needs a tall residual (N >= M) whose panel rows fit the cluster's registers
bool wide_can_block(const DenseBorder& d) {
  static const bool off = std::getenv("QRK_DENSE_UNBLOCKED") != nullptr;     // measurement switch: the BLAS-2 path
  return !off && d.N >= d.M && d.N <= 4LL * kDbCluster * kDbThreads && d.M >= 16;
}

// Q1^T of a banded left block applied to ncols columns, INCLUDING the complement: column j of `src` (n_rows values, leading
// dimension lds) -> column j of d_wx: thin part in rows [0, m1), complement in rows [m1, m1 + w_N)
int banded_left_apply_qt(qrk_solver* h, const double* src, long long lds, double* dst, int ncols) {
  // the buffer between the two phases holds one vector per column
  const size_t groups = (size_t)((h->nb + h->b_group - 1) / h->b_group);
  const size_t gw = groups * ((size_t)(h->b_group - 1) * h->b_step + h->uc);
  if ((size_t)ncols * gw > h->cap_gy) {
    if (h->d_gy) cudaFree(h->d_gy);
    h->d_gy = nullptr; h->cap_gy = 0;
    QRK_TRY_CUDA(h, cudaMalloc(&h->d_gy, (size_t)ncols * gw * sizeof(double)));
    h->cap_gy = (size_t)ncols * gw;
  }
  BandedArgs a = banded_args(h);
  a.b = src; a.ldb = lds;
  a.y = dst; a.ldy = h->w_ld;
  a.comp = dst + h->sum_cols; a.ldcomp = h->w_ld;
  a.ncols = ncols;
  QRK_TRY_CUDA(h, h->bvt->apply_qt(a, h->stream));
  h->launches += banded_launches_per_call();
  return QRK_STATUS_OK;
}

int wide_run_eager(qrk_solver* h, const double* A_in, const double* d_b, double* d_x) {
  const long long n = h->w_ld;
  const int M = h->m2;
  int st;
  if (h->left_banded) {
    st = banded_run(h, A_in, nullptr, nullptr);                         // m_leftSolver.compute (BlockAngularSparseQR.h:472), left = BandedBlockedSparseQR
    if (st != QRK_STATUS_OK) return st;
    if (h->w_ld > h->sum_cols + h->w_N)                                 // the padding row of the leading dimension
      QRK_TRY_CUDA(h, cudaMemsetAsync(h->d_wx, 0, (size_t)h->w_ld * (M + 1) * sizeof(double), h->stream));
    st = banded_left_apply_qt(h, h->d_border, h->ld_border, h->d_wx, M);   // Q1^T J2 (:365)
    if (st != QRK_STATUS_OK) return st;
    if (d_b) {
      st = banded_left_apply_qt(h, d_b, h->n_rows, h->d_wx + (long long)M * n, 1);
      if (st != QRK_STATUS_OK) return st;
    }
  } else {
    st = run_factor(h, A_in, nullptr, nullptr);                         // m_leftSolver.compute (BlockAngularSparseQR.h:472)
    if (st != QRK_STATUS_OK) return st;
    st = run_op(h, OP_APPLY_QT, h->d_border, h->ld_border, h->d_wx, n, M);   // Q1^T J2 (:365)
    if (st != QRK_STATUS_OK) return st;
    if (d_b) {
      st = run_op(h, OP_APPLY_QT, d_b, n, h->d_wx + (long long)M * n, n, 1);
      if (st != QRK_STATUS_OK) return st;
    }
  }
  const DenseBorder d = wide_desc(h, d_b ? 1 : 0);
  h->wide_blocked = false;
  if (d.N > 0) {
    if (d.pivot == 2) {
      st = wide_thin_sparse(h, d);
    } else if (wide_can_block(d)) {
      h->wide_blocked = true;
      st = wide_blocked_qr(h, d);
      if (st == QRK_STATUS_UNSUPPORTED) {   // the panel cluster could not be launched: the column-by-column path instead
        h->wide_blocked = false;
        st = wide_unblocked(h, d);
        if (st != QRK_STATUS_OK) return st;
        return wide_back(h, d_b != nullptr, d_x);
      }
      if (st != QRK_STATUS_OK) return st;
      if (d.pivot) {                      // ColPivHouseholderQR on the triangle: same P2, |R2|, rank and x as on the tall matrix
        dense_extract_tri_kernel<<<148, 256, 0, h->stream>>>(d.A, d.ld, M, M + d.nrhs, h->d_wtri);
        h->launches++;
        const DenseBorder t = wide_tri_desc(h, d.nrhs);
        const size_t smem = tri_colpiv_smem_bytes(M, d.nrhs);
        constexpr size_t kTriMaxDyn = kMaxSmem - 1024;                     // the kernel's static shared memory comes on top
        static const bool no_reg = std::getenv("QRK_TRI_SMEM") != nullptr;     // A/B switch: the shared-memory resident kernel
        if (!no_reg && tri_reg_fits(M, d.nrhs)) {                              // register-resident triangle (dense_tri_reg.cuh)
          dense_tri_colpiv_reg_kernel<<<kDbCluster, kTrThreads, 0, h->stream>>>(t);
          if (cudaGetLastError() == cudaSuccess) h->launches++;
          else st = wide_unblocked(h, t);
        } else if (smem <= kTriMaxDyn) {  // the whole triangle fits the shared memory of one 8-CTA cluster: one launch
          static bool opted[64] = {};
          QRK_TRY_CUDA(h, ensure_smem(dense_tri_colpiv_kernel, kTriMaxDyn, opted));
          dense_tri_colpiv_kernel<<<kDbCluster, kTriThreads, smem, h->stream>>>(t);
          if (cudaGetLastError() == cudaSuccess) h->launches++;
          else st = wide_unblocked(h, t);   // an 8-CTA cluster with this much shared memory could not be placed (partitioned / shared GPU)
        } else {
          st = wide_unblocked(h, t);
        }
      } else {
        dense_prep_kernel<<<1, 256, 0, h->stream>>>(d);                   // P2 = identity (reads the stale norm table only for
        h->launches++;                                                    // the unused threshold)
      }
    } else {
      st = wide_unblocked(h, d);
    }
    if (st != QRK_STATUS_OK) return st;
  }
  return wide_back(h, d_b != nullptr, d_x);
}

// Launch-bound steps are replayed from CUDA graphs.  The wide-border step is ~250 small launches on two streams (panel,
// look-ahead update, trailing update per 8 columns); the narrow-border TSQR step is three dependent launches whose gaps are a
// tenth of the step.  The second time a handle sees the same buffers (values, rhs, x, border, stream) the step is captured --
// a second stream's fork / join and a programmatic dependent launch are captured with it -- and replayed from then on
// (reference test 4 / 5 sizes: 4.0 -> 3.6 ms, 2.0 -> 1.75 ms; config 3: 49.0 -> 46.8 us).  The first call stays eager: it
// makes the lazy allocations and takes the fallbacks (a cluster that cannot be placed) that must not happen inside a capture.
// Not captured: a stream the CALLER is capturing, steps with a host read-back (BlockedThinSparseQR right solver) or that end
// at the per-GPU triangle (NCCL form), QRK_NO_GRAPH=1; any capture failure turns the graph off for the handle.
void step_graph_drop(qrk_solver::StepGraph& g) {
  if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
  g.seen = 0;
}

template <class Eager, class Flag>
int step_graph_run(qrk_solver* h, qrk_solver::StepGraph& g, const void* const (&key)[8], bool allowed, Flag&& flag, Eager&& eager) {
  static const bool no_graph = std::getenv("QRK_NO_GRAPH") != nullptr;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (no_graph || g.off || !allowed || cudaStreamIsCapturing(h->stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)
    return eager();
  const bool same = std::memcmp(key, g.key, sizeof(g.key)) == 0;
  if (same && g.exec) {
    QRK_TRY_CUDA(h, cudaGraphLaunch(g.exec, h->stream));
    h->launches += g.launches;
    flag(g.flag, false);                   // restore the host-side state the eager step leaves behind
    return QRK_STATUS_OK;
  }
  if (!same) { step_graph_drop(g); std::memcpy(g.key, key, sizeof(g.key)); }
  if (g.seen == 0) { g.seen = 1; return eager(); }
  const long long l0 = h->launches;
  auto give_up = [&]() { (void)cudaGetLastError(); g.off = true; step_graph_drop(g); h->launches = l0; return eager(); };
  if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return give_up();
  const int st = eager();
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  if (st != QRK_STATUS_OK || e != cudaSuccess || !graph) { if (graph) cudaGraphDestroy(graph); return give_up(); }
  const cudaError_t ei = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ei != cudaSuccess) { g.exec = nullptr; return give_up(); }
  g.launches = h->launches - l0;
  flag(g.flag, true);                      // record it
  QRK_TRY_CUDA(h, cudaGraphLaunch(g.exec, h->stream));
  return QRK_STATUS_OK;
}

int wide_run(qrk_solver* h, const double* A_in, const double* d_b, double* d_x) {
  const void* key[8] = {A_in, d_b, d_x, h->d_border, reinterpret_cast<const void*>(static_cast<intptr_t>(h->ld_border)), h->stream, nullptr, nullptr};
  return step_graph_run(h, h->sg_wide, key, h->desc.right_solver != QRK_RIGHT_THIN_SPARSE,
                        [&](bool& f, bool record) { if (record) f = h->wide_blocked; else h->wide_blocked = f; },
                        [&]() { return wide_run_eager(h, A_in, d_b, d_x); });
}

int wide_solve_stored(qrk_solver* h, const double* d_b, double* d_x) {
  if (h->thin_deferred) {
    h->err = "BlockedThinSparseQR right solver: the factorisation deferred zero-pivot columns; solve() on the stored factors is not provided, use the fused compute_solve()";
    return QRK_STATUS_UNSUPPORTED;
  }
  const long long n = h->w_ld;
  const int M = h->m2;
  double* rhs_col = h->d_wx + (long long)M * n;
  int st = h->left_banded ? banded_left_apply_qt(h, d_b, h->n_rows, rhs_col, 1) : run_op(h, OP_APPLY_QT, d_b, n, rhs_col, n, 1);
  if (st != QRK_STATUS_OK) return st;
  DenseBorder d = wide_desc(h, 1);
  if (d.N > 0) {
    const bool two_stage = h->wide_blocked && d.pivot == 1;
    if (two_stage) d.tau = h->d_wtau1;
    dense_apply_qt_kernel<1024><<<1, 1024, 0, h->stream>>>(d, rhs_col + h->sum_cols);   // Q2^T on the bottom rows (:619-624)
    h->launches++;
    if (two_stage) {                      // second stage: the ColPiv reflectors of the triangle on the first M entries
      QRK_TRY_CUDA(h, cudaMemcpyAsync(h->d_wtri + (size_t)M * M, rhs_col + h->sum_cols, M * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      const DenseBorder t = wide_tri_desc(h, 1);
      dense_apply_qt_kernel<1024><<<1, 1024, 0, h->stream>>>(t, h->d_wtri + (size_t)M * M);
      h->launches++;
    }
    QRK_TRY_CUDA(h, cudaGetLastError());
  }
  return wide_back(h, true, d_x);
}


// ---- matrixQ() of the block-angular solver: Q = Q1 [I 0; 0 Q2] (BlockAngularSparseQR.h:598-644) ---------------------------
// The dense right block's Q2 acts on the complement rows [m1, n) of the left factor's FullQ layout.
//  * dense right-block path: the stored factorisation (one Householder QR, or the blocked first stage followed by the
//    ColPiv reflectors of its triangle) is applied as it is for solve(b).
//  * fused TSQR path: the tree keeps no reflectors, so Q2 is built on demand from the residual panel kept by compute():
//    an unpivoted Householder QR of Abot in the column order P2 the root chose (R is unique up to row signs for a fixed
//    column order; the signs are aligned with the stored R2 so that Q^T A P_c = R holds entry by entry).
int ensure_q2(qrk_solver* h) {
  if (h->wide || h->q2_ready) return QRK_STATUS_OK;
  QRK_REQUIRE(h, h->have_abot, "matrixQ() of the block-angular solver needs compute(): the fused compute_solve() does not keep the residual panel");
  QRK_REQUIRE(h, h->root_done, "the TSQR root has not run yet (multi-GPU: call qrk_angular_merge first)");
  const long long N = h->n_rows - h->sum_cols;
  const int M = h->m2;
  if (!h->d_q2) {
    QRK_TRY_CUDA(h, cudaMalloc(&h->d_q2, std::max<long long>(1, N * M) * sizeof(double)));
    QRK_TRY_CUDA(h, cudaMalloc(&h->d_q2tau, M * sizeof(double)));
    QRK_TRY_CUDA(h, cudaMalloc(&h->d_q2sign, M * sizeof(double)));
    QRK_TRY_CUDA(h, cudaMalloc(&h->d_q2scr, (2 * M + 2) * sizeof(double)));
    QRK_TRY_CUDA(h, cudaMalloc(&h->d_q2iscr, (M + 1) * sizeof(int)));
  }
  std::vector<int> p2(M);
  QRK_TRY_CUDA(h, cudaMemcpyAsync(p2.data(), h->d_root_i, M * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int c = 0; c < M; c++)
    QRK_TRY_CUDA(h, cudaMemcpyAsync(h->d_q2 + (long long)c * N, h->d_abot + (long long)p2[c] * N, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  QRK_TRY_CUDA(h, cudaMemsetAsync(h->d_q2tau, 0, M * sizeof(double), h->stream));
  DenseBorder d;
  d.A = h->d_q2; d.ld = N; d.N = N; d.Nrule = N; d.M = M; d.nrhs = 0; d.pivot = 0;
  d.upd = h->d_q2scr; d.dir = h->d_q2scr + M; d.scal = h->d_q2scr + 2 * M; d.tau = h->d_q2tau;
  d.perm = h->d_q2iscr; d.iscal = h->d_q2iscr + M;
  if (N > 0) {
    int st = wide_unblocked(h, d);
    if (st != QRK_STATUS_OK) return st;
  }
  dense_diag_sign_kernel<<<1, 64, 0, h->stream>>>(h->d_q2, N, h->d_root, M, (int)std::min<long long>(N, M), h->d_q2sign);
  QRK_TRY_CUDA(h, cudaGetLastError());
  h->launches++;
  h->q2_ready = true;
  return QRK_STATUS_OK;
}

// vec (the n - m1 complement rows of one column) <- Q2^T vec (transpose) or Q2 vec
int apply_q2(qrk_solver* h, double* vec, bool transpose) {
  const int M = h->m2;
  if (!h->wide) {
    DenseBorder d;
    d.A = h->d_q2; d.ld = h->n_rows - h->sum_cols; d.N = d.ld; d.Nrule = d.N; d.M = M; d.nrhs = 0; d.pivot = 0; d.tau = h->d_q2tau;
    d.upd = d.dir = d.scal = nullptr; d.perm = d.iscal = nullptr;
    if (d.N <= 0) return QRK_STATUS_OK;
    if (transpose) {
      dense_apply_qt_kernel<1024><<<1, 1024, 0, h->stream>>>(d, vec);
      dense_scale_head_kernel<<<1, 64, 0, h->stream>>>(vec, h->d_q2sign, (int)std::min<long long>(d.N, M));
    } else {
      dense_scale_head_kernel<<<1, 64, 0, h->stream>>>(vec, h->d_q2sign, (int)std::min<long long>(d.N, M));
      dense_apply_q_kernel<1024><<<1, 1024, 0, h->stream>>>(d, vec);
    }
    h->launches += 2;
    QRK_TRY_CUDA(h, cudaGetLastError());
    return QRK_STATUS_OK;
  }
  DenseBorder d = wide_desc(h, 1);
  if (d.N <= 0) return QRK_STATUS_OK;
  if (h->thin_deferred) {
    h->err = "BlockedThinSparseQR right solver: the factorisation deferred zero-pivot columns; matrixQ() products are not provided for it";
    return QRK_STATUS_UNSUPPORTED;
  }
  const bool two_stage = h->wide_blocked && d.pivot == 1;
  if (two_stage) d.tau = h->d_wtau1;
  double* head = h->d_wtri + (size_t)M * M;        // the right-hand side column of the triangle: scratch for the second stage
  const DenseBorder t = wide_tri_desc(h, 1);
  if (transpose) {
    dense_apply_qt_kernel<1024><<<1, 1024, 0, h->stream>>>(d, vec);
    h->launches++;
    if (two_stage) {
      QRK_TRY_CUDA(h, cudaMemcpyAsync(head, vec, M * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      dense_apply_qt_kernel<1024><<<1, 1024, 0, h->stream>>>(t, head);
      QRK_TRY_CUDA(h, cudaMemcpyAsync(vec, head, M * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      h->launches++;
    }
  } else {
    if (two_stage) {
      QRK_TRY_CUDA(h, cudaMemcpyAsync(head, vec, M * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      dense_apply_q_kernel<1024><<<1, 1024, 0, h->stream>>>(t, head);
      QRK_TRY_CUDA(h, cudaMemcpyAsync(vec, head, M * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      h->launches++;
    }
    dense_apply_q_kernel<1024><<<1, 1024, 0, h->stream>>>(d, vec);
    h->launches++;
  }
  QRK_TRY_CUDA(h, cudaGetLastError());
  return QRK_STATUS_OK;
}

// matrixQ().transpose() * B / matrixQ() * B for a block-angular handle (device pointers, FullQ layout of the left factor)
int angular_apply(qrk_solver* h, int op, const double* d_B, long long ldb, double* d_X, long long ldx, int nrhs) {
  if (h->left_banded) {
    h->err = "matrixQ() products with a banded left solver are not provided (the two-phase banded Q1 has an extended complement)";
    return QRK_STATUS_UNSUPPORTED;
  }
  QRK_REQUIRE(h, h->world == 1, "matrixQ() products of a multi-GPU block-angular handle are not provided (Q2 spans all ranks' rows)");
  int st = ensure_q2(h);
  if (st != QRK_STATUS_OK) return st;
  const long long m1 = h->sum_cols, n = h->n_rows;
  if (op == OP_APPLY_QT) {
    st = run_op(h, OP_APPLY_QT, d_B, ldb, d_X, ldx, nrhs);                       // Q1^T v (:616-618)
    for (int j = 0; j < nrhs && st == QRK_STATUS_OK; j++) st = apply_q2(h, d_X + j * ldx + m1, true);   // Q2^T on rows [m1, n) (:619-624)
    return st;
  }
  // Q v = Q1 [v1; Q2 v2] (:633-642): the bottom rows change before Q1 sees them, so work on a copy
  const long long ldt = (n + 1) & ~1LL;
  st = ensure_buffer(h, h->d_qtmp, h->cap_qtmp, (size_t)ldt * nrhs);
  if (st != QRK_STATUS_OK) return st;
  QRK_TRY_CUDA(h, cudaMemcpy2DAsync(h->d_qtmp, ldt * sizeof(double), d_B, ldb * sizeof(double), n * sizeof(double), nrhs,
                                    cudaMemcpyDeviceToDevice, h->stream));
  for (int j = 0; j < nrhs && st == QRK_STATUS_OK; j++) st = apply_q2(h, h->d_qtmp + j * ldt + m1, false);
  if (st != QRK_STATUS_OK) return st;
  return run_op(h, OP_APPLY_Q, h->d_qtmp, ldt, d_X, ldx, nrhs);
}

int angular_run(qrk_solver* h, const double* A_in, const double* d_b, double* d_x, bool keep_abot) {
  QRK_REQUIRE(h, h->d_border, "no border set: call qrk_set_border first (BlockMatrix1x2 right block)");
  h->q2_ready = false;
  if (h->wide) return wide_run(h, A_in, d_b, d_x);
  if (keep_abot && !h->d_abot)
    QRK_TRY_CUDA(h, cudaMalloc(&h->d_abot, std::max<long long>(1, (h->n_rows - h->sum_cols) * (long long)(h->m2 + 1)) * sizeof(double)));
  auto eager = [&]() -> int {
    AngularArgs a = angular_args(h);
    a.A_in = A_in;
    a.b = d_b;
    if (keep_abot) a.abot = h->d_abot;
    QRK_TRY_CUDA(h, h->avt->factor(a, h->stream));
    h->launches++;
    h->have_abot = keep_abot;
    return angular_root_and_back(h, a, d_b != nullptr, d_x, 0);
  };
  // K1 -> root -> K3 replayed from a CUDA graph (see step_graph_run)
  const void* key[8] = {A_in, d_b, d_x, h->d_border, reinterpret_cast<const void*>(static_cast<intptr_t>(h->ld_border)), h->stream,
                        reinterpret_cast<const void*>(static_cast<intptr_t>(keep_abot ? 1 : 0)),
                        reinterpret_cast<const void*>(static_cast<uintptr_t>(h->xchg_timeout_ns + 1000003ull * (unsigned)(h->xchg_rank + 2)))};
  // (world > 1: also the fused exchange stays eager here -- instantiating a graph can wait for the device, and with several
  //  ranks on ONE device a peer's root kernel may be spinning for this rank's step; callers capture the whole step themselves)
  const bool complete = h->world == 1;
  return step_graph_run(h, h->sg_tsqr, key, complete,
                        [&](bool& f, bool record) { if (record) f = keep_abot; else { h->have_abot = f; h->root_done = true; } }, eager);
}

// solve(b) on a stored factorisation: Q1^T b, TSQR redone over [Abot | b_bot], root, back substitution
int angular_solve_stored(qrk_solver* h, const double* d_b, double* d_x) {
  if (h->wide) return wide_solve_stored(h, d_b, d_x);
  QRK_REQUIRE(h, h->have_abot, "solve() after a fused compute_solve(): the residual panel was not kept; call compute() first");
  auto eager = [&]() -> int {
    AngularArgs a = angular_args(h);
    a.b = d_b;
    a.abot = h->d_abot;
    QRK_TRY_CUDA(h, h->avt->rhs(a, h->stream));
    h->launches++;
    return angular_root_and_back(h, a, true, d_x, 1);
  };
  // the reference's own calling pattern is compute(J) followed by solve(b): the three launches of solve() are replayed from a
  // CUDA graph like those of the fused step (step_graph_run)
  const void* key[8] = {d_b, d_x, h->d_abot, h->d_values, h->stream, nullptr, nullptr, nullptr};
  return step_graph_run(h, h->sg_solve, key, h->world == 1, [&](bool&, bool record) { if (!record) h->root_done = true; }, eager);
}


// ---- banded blocked ----------------------------------------------------------------------------------
BandedArgs banded_args(qrk_solver* h) {
  BandedArgs a;
  a.nb = h->nb; a.packed = h->d_values; a.tau = h->d_btau; a.rband = h->d_rband; a.y = h->d_ythin;
  a.last_cols = (int)(h->sum_cols - (h->nb - 1) * (long long)h->b_step);   // sum_cols = the banded columns (n_cols includes a border)
  a.group = h->b_group; a.gband = h->d_gband; a.gy = h->d_gy; a.cvec = h->d_cvec; a.ctau = h->d_ctau;
  if (h->bgen) a.gen = &h->g_args;
  return a;
}

// window sweep (+ fused Q^T b and back substitution when d_b != nullptr)
int banded_run(qrk_solver* h, const double* A_in, const double* d_b, double* d_x) {
  BandedArgs a = banded_args(h);
  a.A_in = A_in; a.b = d_b;
  QRK_TRY_CUDA(h, h->bvt->factor(a, h->stream));
  h->launches += banded_launches_per_call();
  if (d_b) {
    a.x = d_x;
    QRK_TRY_CUDA(h, h->bvt->backsolve(a, h->stream));
    h->launches++;
  }
  return QRK_STATUS_OK;
}

// matrixR() of the banded solver in the reference's exact compressed layout (BandedBlockedSparseQR.h:484-491, 511-512): window i
// (idxCol, numCols) contributes the dense rectangle rows [idxCol, idxCol + solvedRows) x columns [idxCol, idxCol + numCols) —
// EVERY coefficient of it is a stored entry, the zeros below the diagonal and beyond the band included; solvedRows = the
// distance to the next window's first column, or numRows for the last window (its rows below the triangle are stored zeros
// too).  The windows are the reference's merged blocks (banded_reference_windows).  Values: the band R where it lives, 0 else.
__global__ void export_banded_r_kernel(const double* __restrict__ rband, const int* __restrict__ outer, const int* __restrict__ inner,
                                       double* __restrict__ vals, long long n_cols, long long nb, int bc, int step) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n_cols; j += (long long)gridDim.x * blockDim.x) {
    for (int p = outer[j]; p < outer[j + 1]; p++) {
      const long long g = inner[p];
      double v = 0.0;
      if (g <= j) {                                         // g <= j < n_cols
        long long w = g / step;
        if (w > nb - 1) w = nb - 1;
        const long long off = j - w * step;
        if (off < bc) v = rband[g * bc + off];
      }
      vals[p] = v;
    }
  }
}

// column pointers and row indices of that layout (host; the banded columns only — a border adds its own)
void banded_r_pattern(const qrk_solver* h, std::vector<int>& outer, std::vector<int>* inner) {
  const long long n = h->sum_cols;
  const std::vector<int32_t>& win = h->b_windows;
  const size_t nw = win.size() / 4;
  outer.assign(n + 1, 0);
  auto solved = [&](size_t i) -> long long { return i + 1 < nw ? (long long)win[4 * (i + 1) + 1] - win[4 * i + 1] : (long long)win[4 * i + 2]; };
  for (size_t i = 0; i < nw; i++) {                          // count per column
    const long long c0 = win[4 * i + 1], nc = win[4 * i + 3], sr = solved(i);
    for (long long c = c0; c < c0 + nc && c < n; c++) outer[c + 1] += (int)sr;
  }
  for (long long j = 0; j < n; j++) outer[j + 1] += outer[j];
  if (!inner) return;
  inner->assign((size_t)outer[n], 0);
  std::vector<int> fill(outer.begin(), outer.end() - 1);
  for (size_t i = 0; i < nw; i++) {                          // windows in order: row ranges ascend, so every column stays sorted
    const long long c0 = win[4 * i + 1], nc = win[4 * i + 3], sr = solved(i);
    for (long long c = c0; c < c0 + nc && c < n; c++) {
      int* dst = inner->data() + fill[c];
      for (long long r = 0; r < sr; r++) dst[r] = (int)(c0 + r);
      fill[c] += (int)sr;
    }
  }
}

std::vector<int> banded_r_outer(const qrk_solver* h) {
  std::vector<int> outer;
  banded_r_pattern(h, outer, nullptr);
  return outer;
}

}  // namespace

namespace {

// The general window chain (banded_generic.cuh) from a list of dense blocks {idxRow, idxCol, numRows, numCols}: rows contiguous
// and in order, first columns non-decreasing, every window tall enough to finalise the rows the next one starts after.
// in_cols[i]: columns of block i as stored in the input values (block i at voff[i], column-major numRows x in_cols).
int build_generic_chain(qrk_solver* h, const std::vector<int32_t>& b4, const std::vector<long long>& voff, const std::vector<int>& in_cols) {
  const size_t nw = b4.size() / 4;
  h->g_win.assign(nw, GenWindow());
  long long row = 0, poff = 0, toff = 0, coff = 0;
  int carry = 0, prev_end = 0, max_rows = 1, max_cols = 1;
  for (size_t i = 0; i < nw; i++) {
    GenWindow& w = h->g_win[i];
    w.row0 = b4[4 * i]; w.col0 = b4[4 * i + 1]; w.nrows = b4[4 * i + 2];
    const int ncols_blk = b4[4 * i + 3];
    if (w.row0 != row || w.nrows <= 0 || ncols_blk <= 0 || w.col0 < 0 || (i > 0 && w.col0 < h->g_win[i - 1].col0)) {
      h->err = "banded window chain: blocks must cover the rows contiguously, in order, with non-decreasing first columns";
      return QRK_STATUS_INVALID_ARGUMENT;
    }
    if (i > 0 && w.col0 > prev_end) { h->err = "banded window chain: a column range is covered by no block"; return QRK_STATUS_INVALID_ARGUMENT; }
    w.ncols_in = in_cols[i];
    w.ncols = std::max(ncols_blk, prev_end - w.col0);      // wide enough for the carried rows' columns (:502)
    w.carry = carry;
    const int rows = carry + w.nrows;
    w.steps = std::min(rows, w.ncols);
    const bool last = i + 1 == nw;
    w.solved = last ? w.steps : b4[4 * (i + 1) + 1] - w.col0;
    if (w.solved > w.steps || (last && (w.steps != w.ncols || w.col0 + w.ncols != h->sum_cols))) {
      h->err = "banded window chain: a window has fewer rows than the columns it must finalise (structurally rank deficient)";
      return QRK_STATUS_INVALID_ARGUMENT;
    }
    w.voff = voff[i]; w.poff = poff; w.toff = toff; w.coff = coff;
    poff += (long long)rows * w.ncols; toff += w.steps; coff += rows - w.steps;
    carry = w.steps - w.solved;
    prev_end = w.col0 + w.ncols;
    row += w.nrows;
    max_rows = std::max(max_rows, rows); max_cols = std::max(max_cols, w.ncols);
  }
  if (row != h->n_rows || coff != h->n_rows - h->sum_cols) { h->err = "banded window chain: the blocks do not add up to the matrix"; return QRK_STATUS_INVALID_ARGUMENT; }
  if (banded_generic_smem_bytes(max_rows, max_cols) > kMaxSmem - 2048) {
    h->err = "banded window chain: a window does not fit the shared memory of one SM";
    return QRK_STATUS_UNSUPPORTED;
  }
  std::vector<int> col0(nw + 1);
  for (size_t i = 0; i < nw; i++) col0[i] = h->g_win[i].col0;
  col0[nw] = (int)h->sum_cols;
  QRK_TRY_CUDA(h, cudaMalloc(&h->d_gwin, nw * sizeof(GenWindow)));
  QRK_TRY_CUDA(h, cudaMalloc(&h->d_gcol0, (nw + 1) * sizeof(int)));
  QRK_TRY_CUDA(h, cudaMalloc(&h->d_gpacked, std::max<long long>(poff, 1) * sizeof(double)));
  QRK_TRY_CUDA(h, cudaMalloc(&h->d_gtau, std::max<long long>(toff, 1) * sizeof(double)));
  QRK_TRY_CUDA(h, cudaMalloc(&h->d_gcomp, std::max<long long>(coff, 1) * sizeof(double)));
  QRK_TRY_CUDA(h, cudaMemcpy(h->d_gwin, h->g_win.data(), nw * sizeof(GenWindow), cudaMemcpyHostToDevice));
  QRK_TRY_CUDA(h, cudaMemcpy(h->d_gcol0, col0.data(), (nw + 1) * sizeof(int), cudaMemcpyHostToDevice));
  GenArgs& g = h->g_args;
  g.win = h->d_gwin; g.nwin = (int)nw; g.max_rows = max_rows; g.max_cols = max_cols;
  g.packed = h->d_gpacked; g.tau = h->d_gtau; g.n_rows = h->n_rows; g.n_cols = h->sum_cols;
  h->bgen = true;
  h->bvt = banded_generic_vtable();
  return QRK_STATUS_OK;
}

int peer_exchange_status(qrk_solver* h) {
  if (!h->d_xchg_err || h->xchg_rank < 0) return QRK_STATUS_OK;
  int flag = 0;
  if (cudaMemcpy(&flag, h->d_xchg_err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) { (void)cudaGetLastError(); return QRK_STATUS_CUDA_ERROR; }
  if (flag) { h->err = "fused peer exchange: a peer never delivered its triangle within the timeout; results are NaN-poisoned until the next qrk_angular_p2p_attach"; return QRK_STATUS_PEER_TIMEOUT; }
  return QRK_STATUS_OK;
}

}  // namespace

extern "C" {

int qrk_version(void) { return 100; }
#ifdef QRK_TRI_TRACE
__attribute__((visibility("default"))) int qrk_debug_tri_trace(long long* out16) {
  return cudaMemcpyFromSymbol(out16, qrk::g_tri_trace, sizeof(long long) * 16) == cudaSuccess ? 0 : 1;
}
__attribute__((visibility("default"))) int qrk_debug_panel_trace(long long* out16) {
  return cudaMemcpyFromSymbol(out16, qrk::g_panel_trace, sizeof(long long) * 16) == cudaSuccess ? 0 : 1;
}
#endif

const char* qrk_status_string(int status) {
  switch (status) {
    case QRK_STATUS_OK: return "ok";
    case QRK_STATUS_INVALID_ARGUMENT: return "invalid argument";
    case QRK_STATUS_NOT_FACTORIZED: return "the factorization should be called first, use compute()";
    case QRK_STATUS_CUDA_ERROR: return "CUDA error";
    case QRK_STATUS_NO_DEVICE: return "no CUDA device (qrkit_b200 has no CPU fallback)";
    case QRK_STATUS_ALLOC_FAILED: return "device allocation failed";
    case QRK_STATUS_UNSUPPORTED: return "unsupported configuration";
    case QRK_STATUS_PEER_TIMEOUT: return "fused peer exchange timed out: a peer GPU never delivered its triangle (results are NaN)";
    default: return "unknown status";
  }
}

const char* qrk_last_error(qrk_handle_t h) { return h ? h->err.c_str() : ""; }

int qrk_device_count(int* count) {
  if (!count) return QRK_STATUS_INVALID_ARGUMENT;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
  *count = n;
  return QRK_STATUS_OK;
}

static int create_impl(const qrk_desc_t* desc, const int32_t* gen_blocks, qrk_handle_t* out);

int qrk_create(const qrk_desc_t* desc, qrk_handle_t* out) { return create_impl(desc, nullptr, out); }

int qrk_create_banded_general(const int32_t* blocks, int64_t num_blocks, int64_t n_rows, int64_t n_cols, int32_t device,
                              int32_t suggested_block_cols, qrk_handle_t* out) {
  if (!blocks || num_blocks < 1 || n_rows < 1 || n_cols < 1 || !out) return QRK_STATUS_INVALID_ARGUMENT;
  std::vector<int32_t> rows((size_t)num_blocks), cols((size_t)num_blocks);
  for (int64_t i = 0; i < num_blocks; i++) { rows[(size_t)i] = blocks[4 * i + 2]; cols[(size_t)i] = blocks[4 * i + 3]; }
  qrk_desc_t d = qrk_desc_t();
  d.kind = QRK_BANDED_BLOCKED; d.device = device; d.num_blocks = num_blocks; d.rows = rows.data(); d.cols = cols.data();
  d.n_rows = n_rows; d.n_cols = n_cols; d.reserved[0] = suggested_block_cols;
  return create_impl(&d, blocks, out);
}

static int create_impl(const qrk_desc_t* desc, const int32_t* gen_blocks, qrk_handle_t* out) {
  if (!desc || !out) return QRK_STATUS_INVALID_ARGUMENT;
  *out = nullptr;
  if (desc->kind != QRK_BLOCK_DIAGONAL && desc->kind != QRK_BLOCK_ANGULAR && desc->kind != QRK_BANDED_BLOCKED) return QRK_STATUS_UNSUPPORTED;
  const bool angular = desc->kind == QRK_BLOCK_ANGULAR;
  const bool left_banded = angular && desc->left_solver == QRK_LEFT_BANDED_BLOCKED;
  const bool banded = desc->kind == QRK_BANDED_BLOCKED || left_banded;   // the slab layout / banded kernels of J (or J1)
  if (desc->num_blocks < 0) return QRK_STATUS_INVALID_ARGUMENT;
  const bool uniform = desc->block_rows > 0 && desc->block_cols > 0;
  if (!uniform && desc->num_blocks > 0 && (!desc->rows || !desc->cols)) return QRK_STATUS_INVALID_ARGUMENT;
  int ndev = 0;
  qrk_device_count(&ndev);
  if (ndev <= 0) return QRK_STATUS_NO_DEVICE;
  if (desc->device < 0 || desc->device >= ndev) return QRK_STATUS_INVALID_ARGUMENT;

  qrk_solver* h = new (std::nothrow) qrk_solver;
  if (!h) return QRK_STATUS_ALLOC_FAILED;
  h->desc = *desc;
  h->desc.rows = h->desc.cols = nullptr;
  h->device = desc->device;
  h->nb = desc->num_blocks;
  h->uniform = uniform;
  DeviceGuard g(h->device);
  auto fail = [&](int st) { free_dev(h); if (h->own_stream) cudaStreamDestroy(h->own_stream); delete h; return st; };

  // prefix sums = block origins (BlockDiagonalSparseQR.h:524-525)
  bool landscape = false;
  if (uniform) {
    h->ur = desc->block_rows; h->uc = desc->block_cols;
    h->max_r = h->ur; h->max_c = h->uc;
    h->sum_rows = h->nb * h->ur; h->sum_cols = h->nb * h->uc;
    h->total_values = h->nb * (long long)h->ur * h->uc;
    landscape = h->ur < h->uc;
  } else {
    h->h_rows.assign(desc->rows, desc->rows + h->nb);
    h->h_cols.assign(desc->cols, desc->cols + h->nb);
    h->h_voff.resize(h->nb); h->h_roff.resize(h->nb); h->h_coff.resize(h->nb);
    long long vo = 0, ro = 0, co = 0;
    for (long long i = 0; i < h->nb; i++) {
      const int r = h->h_rows[i], c = h->h_cols[i];
      if (r <= 0 || c <= 0) return fail(QRK_STATUS_INVALID_ARGUMENT);
      if (r < c) landscape = true;
      h->h_voff[i] = vo; h->h_roff[i] = ro; h->h_coff[i] = co;
      vo += (long long)r * c; ro += r; co += c;
      h->max_r = std::max(h->max_r, r); h->max_c = std::max(h->max_c, c);
    }
    h->total_values = vo; h->sum_rows = ro; h->sum_cols = co;
  }
  h->n_rows = desc->n_rows > 0 ? desc->n_rows : h->sum_rows;
  h->n_cols = desc->n_cols > 0 ? desc->n_cols : h->sum_cols;
  if (banded && gen_blocks) {
    // a general chain of dense blocks (BandedBlockedSparseQR.h:408-426 after row ordering and block detection)
    if (desc->reserved[0] < 0 || desc->n_rows != h->sum_rows || desc->n_cols < 1 || desc->n_rows < desc->n_cols) return fail(QRK_STATUS_INVALID_ARGUMENT);
    h->n_rows = desc->n_rows; h->n_cols = desc->n_cols; h->sum_cols = h->n_cols;
    h->info = QRK_INFO_SUCCESS;
    DeviceGuard gg(h->device);
    std::vector<int32_t> b4(gen_blocks, gen_blocks + 4 * h->nb);
    banded_reference_windows_from_blocks(b4, desc->reserved[0] > 0 ? desc->reserved[0] : 2, h->b_windows);
    const int stg = build_generic_chain(h, b4, h->h_voff, h->h_cols);
    if (stg != QRK_STATUS_OK) return fail(stg);
  } else if (banded) {
    // nb block rows of block_rows x block_cols, consecutive blocks shifted by S = block_cols - overlap columns
    // (BlockBandedMatrixInfo::fromBlockBandedPattern, SparseQRUtils.h:274-302)
    if (!uniform || h->nb < 1) return fail(QRK_STATUS_INVALID_ARGUMENT);
    h->b_ov = desc->block_overlap;
    h->b_step = h->uc - h->b_ov;
    if (h->b_step <= 0 || h->b_ov < 0) return fail(QRK_STATUS_INVALID_ARGUMENT);
    h->bvt = banded_vtable(h->ur, h->uc, h->b_ov);
    if (std::getenv("QRK_BANDED_GENERIC")) h->bvt = nullptr;            // test switch: the general window chain for every shape
    if (!h->bvt && left_banded) return fail(QRK_STATUS_UNSUPPORTED);    // the block-angular combination needs an instantiated shape
    if (desc->reserved[0] < 0) return fail(QRK_STATUS_INVALID_ARGUMENT);
    h->n_rows = h->sum_rows;
    // n_cols defaults to the full width of the last slab; a narrower last slab (the reference's pattern gives the last
    // block block_cols - overlap columns, SparseQRUtils.h:284) is selected by passing n_cols explicitly
    const long long full = (h->nb - 1) * (long long)h->b_step + h->uc;
    h->n_cols = desc->n_cols > 0 ? desc->n_cols - (left_banded ? desc->border_cols : 0) : full;   // block angular: desc->n_cols = m1 + m2
    if (h->n_cols > full || h->n_cols <= (h->nb - 1) * (long long)h->b_step) return fail(QRK_STATUS_INVALID_ARGUMENT);
    h->sum_cols = h->n_cols;
    if (h->n_rows < h->n_cols) return fail(QRK_STATUS_INVALID_ARGUMENT);
    // the reference's windows for this geometry: they fix the stored pattern of matrixR() (explicit zeros included)
    banded_reference_windows(h->nb, h->ur, h->uc, h->b_ov, (int)(h->n_cols - (h->nb - 1) * (long long)h->b_step),
                             desc->reserved[0] > 0 ? desc->reserved[0] : 2, h->b_windows);
    if (desc->n_rows > 0 && desc->n_rows != h->n_rows) return fail(QRK_STATUS_INVALID_ARGUMENT);
    h->info = QRK_INFO_SUCCESS;      // landscape slabs are normal here
    if (!h->bvt) {                   // a slab shape that is not instantiated: the general window chain, one window per slab
      DeviceGuard gg(h->device);
      std::vector<int32_t> b4((size_t)h->nb * 4);
      std::vector<long long> voff((size_t)h->nb);
      std::vector<int> in_cols((size_t)h->nb);
      const int last_cols = (int)(h->n_cols - (h->nb - 1) * (long long)h->b_step);
      for (long long k = 0; k < h->nb; k++) {
        b4[4 * k] = (int32_t)(k * h->ur); b4[4 * k + 1] = (int32_t)(k * h->b_step); b4[4 * k + 2] = h->ur;
        b4[4 * k + 3] = (k == h->nb - 1) ? last_cols : h->uc;
        voff[(size_t)k] = k * (long long)h->ur * h->uc; in_cols[(size_t)k] = b4[4 * k + 3];
      }
      const int stg = build_generic_chain(h, b4, voff, in_cols);
      if (stg != QRK_STATUS_OK) return fail(stg);
    }
  }
  if (angular) {
    // left block: uniform small blocks covering all rows, FullQ; border: 1..8 dense columns
    h->m2 = desc->border_cols;
    h->avt = angular_vtable(h->m2);
    if (desc->q_format != QRK_FULL_Q || h->m2 < 1 || h->m2 > 4096) return fail(QRK_STATUS_UNSUPPORTED);
    if (desc->right_solver != QRK_RIGHT_COLPIV && desc->right_solver != QRK_RIGHT_UNPIVOTED && desc->right_solver != QRK_RIGHT_THIN_SPARSE) return fail(QRK_STATUS_INVALID_ARGUMENT);
    if (desc->reserved[1] < 0) return fail(QRK_STATUS_INVALID_ARGUMENT);
    h->left_banded = left_banded;
    if (!uniform || !h->avt || left_banded || !h->avt->shape_ok(h->ur, h->uc) || desc->right_solver != QRK_RIGHT_COLPIV) {   // dense_border.cuh
      h->avt = nullptr;
      h->wide = true;
    }
    if (h->n_rows != h->sum_rows || (desc->n_cols > 0 && desc->n_cols != h->sum_cols + h->m2)) return fail(QRK_STATUS_INVALID_ARGUMENT);
    h->n_cols = h->sum_cols + h->m2;     // cols() = m1 + m2
    h->w_ld = h->n_rows; h->w_N = h->n_rows - h->sum_cols;
  }
  if (h->n_rows < h->sum_rows || h->n_cols < h->sum_cols) return fail(QRK_STATUS_INVALID_ARGUMENT);
  if (h->n_rows > INT32_MAX || h->n_cols > INT32_MAX) return fail(QRK_STATUS_UNSUPPORTED);  // StorageIndex = int
  if (desc->q_format != QRK_FULL_Q && desc->q_format != QRK_BLOCK_DIAGONAL_Q) h->info = QRK_INFO_INVALID_INPUT;  // :501-505
  if (landscape && !banded) h->info = QRK_INFO_INVALID_INPUT;                                                               // :509-516

  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(QRK_STATUS_CUDA_ERROR);
  h->stream = h->own_stream;

  h->small_path = uniform && !banded && small_shape_available(h->ur, h->uc);
  auto up = [&](auto*& dptr, const auto& vec) -> cudaError_t {
    using T = typename std::remove_reference<decltype(vec[0])>::type;
    cudaError_t e = cudaMalloc(&dptr, std::max<size_t>(vec.size(), 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(dptr, vec.data(), vec.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream);
  };
  if (!uniform && h->nb > 0) {
    if (up(h->d_rows, h->h_rows) != cudaSuccess || up(h->d_cols, h->h_cols) != cudaSuccess ||
        up(h->d_voff, h->h_voff) != cudaSuccess || up(h->d_roff, h->h_roff) != cudaSuccess ||
        up(h->d_coff, h->h_coff) != cudaSuccess)
      return fail(QRK_STATUS_ALLOC_FAILED);
  }
  if (!h->small_path && !banded && h->nb > 0 && h->info == QRK_INFO_SUCCESS) {
    // size classes of the generic kernel: (team warps, shared memory rounded up to 8 KB steps)
    const bool piv_cfg = desc->pivoting == QRK_PIVOT_COLPIV;
    if (uniform) {
      SizeClass sc;
      if (wy_eligible(h->ur, h->uc, piv_cfg)) {
        sc.wy_mr = wy_mr(h->ur, h->uc);
        sc.warps = wy_warps(h->ur, h->uc);
        sc.smem = wy_smem_bytes(h->ur, h->uc);
      } else {
        sc.warps = team_warps_for(h->ur, h->uc);
        sc.smem = generic_smem_bytes(h->ur, h->uc);
      }
      sc.count = h->nb;
      if (sc.smem > kMaxSmem) return fail(QRK_STATUS_UNSUPPORTED);
      h->classes.push_back(sc);
    } else {
      // key: (shared memory class, kernel kind (0 generic / MR of the WY kernel), team warps)
      std::map<std::tuple<size_t, int, int>, std::vector<int>> bins;
      for (long long i = 0; i < h->nb; i++) {
        const int r = h->h_rows[i], c = h->h_cols[i];
        if (wy_eligible(r, c, piv_cfg)) {
          const size_t cls = std::min(kMaxSmem, (wy_smem_bytes(r, c) + 2047) / 2048 * 2048);
          bins[{cls, wy_mr(r, c), wy_warps(r, c)}].push_back((int)i);
        } else {
          const size_t need = generic_smem_bytes(r, c);
          if (need > kMaxSmem) return fail(QRK_STATUS_UNSUPPORTED);
          const size_t cls = std::min(kMaxSmem, (need + 8191) / 8192 * 8192);
          bins[{cls, 0, team_warps_for(r, c)}].push_back((int)i);
        }
      }
      for (auto it = bins.rbegin(); it != bins.rend(); ++it) {   // largest first: long CTAs start early
        SizeClass sc;
        sc.smem = std::get<0>(it->first); sc.wy_mr = std::get<1>(it->first); sc.warps = std::get<2>(it->first);
        sc.count = (long long)it->second.size();
        if (up(sc.d_ids, it->second) != cudaSuccess) return fail(QRK_STATUS_ALLOC_FAILED);
        h->classes.push_back(sc);
      }
    }
  }
  if (cudaMalloc(&h->d_tau, std::max<long long>(h->n_cols, 1) * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&h->d_perm, std::max<long long>(h->n_cols, 1) * sizeof(int)) != cudaSuccess)
    return fail(QRK_STATUS_ALLOC_FAILED);
  cudaMemsetAsync(h->d_tau, 0, std::max<long long>(h->n_cols, 1) * sizeof(double), h->stream);
  iota_kernel<<<256, 256, 0, h->stream>>>(h->d_perm, h->n_cols);   // m_outputPerm_c.setIdentity (:417)
  h->launches++;
  if (banded && h->bgen) {
    if (cudaMalloc(&h->d_ythin, (size_t)h->sum_cols * sizeof(double)) != cudaSuccess) return fail(QRK_STATUS_ALLOC_FAILED);
  } else if (banded) {
    // slabs per parallel group (banded.cuh): enough groups to fill the GPU with one warp each, few enough that the
    // overlap rows the chase re-eliminates at every group boundary stay a small fraction (OV per group * S columns)
    h->b_group = (int)std::max<long long>(8, std::min<long long>(64, h->nb / 2048));
    if (const char* env = std::getenv("QRK_BANDED_GROUP")) h->b_group = std::max(1, std::atoi(env));
    const size_t groups = (size_t)((h->nb + h->b_group - 1) / h->b_group);
    const size_t gw = groups * ((size_t)(h->b_group - 1) * h->b_step + h->uc);
    if (left_banded) {
      const int last_cols = (int)(h->sum_cols - (h->nb - 1) * (long long)h->b_step);
      h->w_N = banded_comp_rows(h->nb, h->ur, h->uc, h->b_ov, h->b_group, last_cols);
      h->w_ld = (h->sum_cols + h->w_N + 1) & ~1LL;
    }
    if (cudaMalloc(&h->d_rband, (size_t)h->sum_cols * h->uc * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_btau, (size_t)h->nb * h->uc * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_ythin, (size_t)h->sum_cols * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_gband, gw * h->uc * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_gy, gw * sizeof(double)) != cudaSuccess || (h->cap_gy = gw) == 0 ||
        cudaMalloc(&h->d_cvec, std::max<size_t>(1, gw * h->b_ov) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_ctau, 2 * gw * sizeof(double)) != cudaSuccess)
      return fail(QRK_STATUS_ALLOC_FAILED);
  }
  if (angular && h->wide) {
    const size_t M = (size_t)h->m2;
    if (cudaMalloc(&h->d_wx, std::max<size_t>(1, (size_t)h->w_ld * (M + 1)) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_wupd, M * sizeof(double)) != cudaSuccess || cudaMalloc(&h->d_wdir, M * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_wtau2, M * sizeof(double)) != cudaSuccess || cudaMalloc(&h->d_wscal, 2 * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_wperm, M * sizeof(int)) != cudaSuccess || cudaMalloc(&h->d_wiscal, 2 * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&h->d_wtau1, M * sizeof(double)) != cudaSuccess || cudaMalloc(&h->d_wT, 64 * ((M + 7) / 8) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_wtri, M * (M + 1) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_wpart, ((M + 1 + 7) / 8 + 1) * (size_t)kWideRowSplit * 64 * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_root, (M * M + 3 * M) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_root_i, (M + 1) * sizeof(int)) != cudaSuccess)
      return fail(QRK_STATUS_ALLOC_FAILED);
    cudaMemsetAsync(h->d_root, 0, (M * M + 3 * M) * sizeof(double), h->stream);
    cudaMemsetAsync(h->d_root_i, 0, (M + 1) * sizeof(int), h->stream);
    cudaMemsetAsync(h->d_wtau2, 0, M * sizeof(double), h->stream);
    cudaMemsetAsync(h->d_wupd, 0, M * sizeof(double), h->stream);     // norm tables: read (and ignored: 0 = no downdate) before a
    cudaMemsetAsync(h->d_wdir, 0, M * sizeof(double), h->stream);     // panel / path has computed its norms (initcheck-clean)
  } else if (angular) {
    const bool piv = desc->pivoting == QRK_PIVOT_COLPIV;
    if (h->avt->max_grid(h->ur, h->uc, piv, &h->a_grid) != cudaSuccess) return fail(QRK_STATUS_CUDA_ERROR);
    const size_t tri = (size_t)h->avt->tri_doubles;
    if (cudaMalloc(&h->d_partials, std::max<size_t>(1, (size_t)h->a_grid * tri) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_tri, tri * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_root, (size_t)(h->m2 * h->m2 + 3 * h->m2) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_root_i, (size_t)(h->m2 + 1) * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&h->d_atop, std::max<long long>(1, h->sum_cols * (long long)h->m2) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_y1, std::max<long long>(1, h->sum_cols) * sizeof(double)) != cudaSuccess)
      return fail(QRK_STATUS_ALLOC_FAILED);
    cudaMemsetAsync(h->d_root, 0, (size_t)(h->m2 * h->m2 + 3 * h->m2) * sizeof(double), h->stream);
    cudaMemsetAsync(h->d_root_i, 0, (size_t)(h->m2 + 1) * sizeof(int), h->stream);
  }
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return fail(QRK_STATUS_CUDA_ERROR);  // host vectors may go away
  *out = h;
  return QRK_STATUS_OK;
}

int qrk_destroy(qrk_handle_t h) {
  if (!h) return QRK_STATUS_OK;
  {
    DeviceGuard g(h->device);
    cudaStreamSynchronize(h->stream);
    free_dev(h);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
  }
  delete h;
  return QRK_STATUS_OK;
}

int qrk_set_stream(qrk_handle_t h, void* cuda_stream) {
  if (!h) return QRK_STATUS_INVALID_ARGUMENT;
  h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  for (auto* g : {&h->sg_wide, &h->sg_tsqr, &h->sg_solve}) { if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; } g->seen = 0; }
  return QRK_STATUS_OK;
}

int qrk_synchronize(qrk_handle_t h) {
  if (!h) return QRK_STATUS_INVALID_ARGUMENT;
  DeviceGuard g(h->device);
  QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
  return peer_exchange_status(h);
}

int qrk_total_values(qrk_handle_t h, int64_t* n) {
  if (!h || !n) return QRK_STATUS_INVALID_ARGUMENT;
  *n = h->total_values;
  return QRK_STATUS_OK;
}

// ---- NUMA-local pinned host memory (Linux; silently a no-op where sysfs / the syscalls are unavailable) ----------------
static int parse_cpulist(const char* s, cpu_set_t* set) {
  CPU_ZERO(set);
  int count = 0;
  while (*s) {
    char* end = nullptr;
    long a = std::strtol(s, &end, 10);
    if (end == s) break;
    long b = a;
    s = end;
    if (*s == '-') { b = std::strtol(s + 1, &end, 10); s = end; }
    for (long c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET((int)c, set); count++; }
    if (*s == ',') s++; else break;
  }
  return count;
}

int qrk_bind_host_thread_to_device(int32_t device, int32_t* numa_node, int32_t* cpus_bound) {
  if (numa_node) *numa_node = -1;
  if (cpus_bound) *cpus_bound = 0;
  int ndev = 0;
  qrk_device_count(&ndev);
  if (ndev <= 0) return QRK_STATUS_NO_DEVICE;
  if (device < 0 || device >= ndev) return QRK_STATUS_INVALID_ARGUMENT;
  char bus[64] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus) - 1, device) != cudaSuccess) { (void)cudaGetLastError(); return QRK_STATUS_OK; }
  for (char* p = bus; *p; ++p) *p = (char)std::tolower((unsigned char)*p);
  char path[160];
  std::snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  int node = -1;
  if (FILE* f = std::fopen(path, "r")) { if (std::fscanf(f, "%d", &node) != 1) node = -1; std::fclose(f); }
  std::snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
  char list[4096] = {0};
  if (FILE* f = std::fopen(path, "r")) { if (!std::fgets(list, sizeof(list) - 1, f)) list[0] = 0; std::fclose(f); }
  cpu_set_t want, allowed, both;
  const int n_local = parse_cpulist(list, &want);
  int bound = 0;
  if (n_local > 0 && sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
    CPU_AND(&both, &want, &allowed);                       // never leave the cpuset the container grants
    bound = CPU_COUNT(&both);
    if (bound > 0 && sched_setaffinity(0, sizeof(both), &both) != 0) bound = 0;
  }
  if (node >= 0 && node < 1024) {                          // MPOL_PREFERRED: allocations of this thread prefer the GPU's node
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1UL << (node % (8 * sizeof(unsigned long)));
    (void)syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, (unsigned long)(8 * sizeof(mask)));
  }
  if (numa_node) *numa_node = node;
  if (cpus_bound) *cpus_bound = bound;
  return QRK_STATUS_OK;
}

int qrk_host_alloc(void** ptr, int64_t bytes, int32_t device) {
  if (!ptr || bytes <= 0) return QRK_STATUS_INVALID_ARGUMENT;
  *ptr = nullptr;
  int st = qrk_bind_host_thread_to_device(device, nullptr, nullptr);
  if (st != QRK_STATUS_OK) return st;
  DeviceGuard g(device);
  void* p = nullptr;
  if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); return QRK_STATUS_ALLOC_FAILED; }
  // pages are pinned at allocation; touching them here keeps the placement decision inside this (bound) thread even
  // on drivers that populate lazily
  volatile char* c = static_cast<volatile char*>(p);
  for (int64_t i = 0; i < bytes; i += 4096) c[i] = 0;
  *ptr = p;
  return QRK_STATUS_OK;
}

int qrk_host_free(void* ptr) {
  if (!ptr) return QRK_STATUS_OK;
  return cudaFreeHost(ptr) == cudaSuccess ? QRK_STATUS_OK : QRK_STATUS_CUDA_ERROR;
}

int qrk_set_blocks(qrk_handle_t h, const double* values, int memspace) {
  if (!h) return QRK_STATUS_INVALID_ARGUMENT;
  QRK_REQUIRE(h, values || h->total_values == 0, "values is null");
  DeviceGuard g(h->device);
  {
    int st = ensure_own_values(h);
    if (st != QRK_STATUS_OK) return st;
  }
  QRK_TRY_CUDA(h, cudaMemcpyAsync(h->d_values, values, h->total_values * sizeof(double),
                                  memspace == QRK_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  h->has_blocks = true;
  h->factorized = false;
  return QRK_STATUS_OK;
}

int qrk_adopt_blocks(qrk_handle_t h, double* device_values) {
  if (!h) return QRK_STATUS_INVALID_ARGUMENT;
  QRK_REQUIRE(h, device_values || h->total_values == 0, "device_values is null");
  QRK_REQUIRE(h, aligned16(device_values), "device_values must be 16-byte aligned");
  DeviceGuard g(h->device);
  if (h->own_values && h->d_values) cudaFree(h->d_values);
  h->d_values = device_values;
  h->own_values = false;
  h->has_blocks = true;
  h->factorized = false;
  return QRK_STATUS_OK;
}

int qrk_analyze_pattern(qrk_handle_t h, const int32_t* row_perm) {
  if (!h) return QRK_STATUS_INVALID_ARGUMENT;
  h->h_rowperm.resize(h->n_rows);
  if (row_perm) std::copy(row_perm, row_perm + h->n_rows, h->h_rowperm.begin());
  else std::iota(h->h_rowperm.begin(), h->h_rowperm.end(), 0);
  h->analyzed = true;
  return QRK_STATUS_OK;
}

int qrk_factorize(qrk_handle_t h) {
  if (!h) return QRK_STATUS_INVALID_ARGUMENT;
  QRK_REQUIRE(h, h->has_blocks, "no blocks set: call qrk_set_blocks / qrk_adopt_blocks first");
  if (!h->analyzed) qrk_analyze_pattern(h, nullptr);
  if (h->info == QRK_INFO_INVALID_INPUT) return QRK_STATUS_OK;   // reported through info(), as the reference
  DeviceGuard g(h->device);
  int st = ang(h) ? angular_run(h, h->d_values, nullptr, nullptr, true)
                  : (h->bvt ? banded_run(h, h->d_values, nullptr, nullptr) : run_factor(h, h->d_values, nullptr, nullptr));
  if (st != QRK_STATUS_OK) return st;
  h->factorized = true;
  return QRK_STATUS_OK;
}

// Device-resident input: the kernels read the caller's blocks directly and write the packed factors
// into the handle's storage (the input is borrowed const for the call, as compute(const MatrixType&)).
static int compute_from_device(qrk_solver* h, const double* values, const double* d_b, double* d_x) {
  QRK_REQUIRE(h, values || h->total_values == 0, "values is null");
  QRK_REQUIRE(h, aligned16(values), "device values must be 16-byte aligned");
  DeviceGuard g(h->device);
  int st = ensure_own_values(h);
  if (st != QRK_STATUS_OK) return st;
  h->has_blocks = true;
  h->factorized = false;
  if (!h->analyzed) qrk_analyze_pattern(h, nullptr);
  if (h->info == QRK_INFO_INVALID_INPUT) return QRK_STATUS_OK;
  if (d_x && h->n_cols > h->sum_cols && !ang(h))
    QRK_TRY_CUDA(h, cudaMemsetAsync(d_x + h->sum_cols, 0, (h->n_cols - h->sum_cols) * sizeof(double), h->stream));
  st = ang(h) ? angular_run(h, values, d_b, d_x, d_b == nullptr) : (h->bvt ? banded_run(h, values, d_b, d_x) : run_factor(h, values, d_b, d_x));
  if (st != QRK_STATUS_OK) return st;
  h->factorized = true;
  return QRK_STATUS_OK;
}

int qrk_compute(qrk_handle_t h, const double* values, int memspace) {
  if (!h) return QRK_STATUS_INVALID_ARGUMENT;
  if (memspace == QRK_DEVICE) return compute_from_device(h, values, nullptr, nullptr);
  int st = qrk_set_blocks(h, values, memspace);
  if (st != QRK_STATUS_OK) return st;
  st = qrk_analyze_pattern(h, nullptr);
  if (st != QRK_STATUS_OK) return st;
  return qrk_factorize(h);
}

int qrk_factorize_solve(qrk_handle_t h, const double* b, double* x, int memspace) {
  if (!h) return QRK_STATUS_INVALID_ARGUMENT;
  QRK_REQUIRE(h, h->has_blocks, "no blocks set: call qrk_set_blocks / qrk_adopt_blocks first");
  QRK_REQUIRE(h, b && x, "b / x is null");
  if (!h->analyzed) qrk_analyze_pattern(h, nullptr);
  if (h->info == QRK_INFO_INVALID_INPUT) return QRK_STATUS_OK;
  DeviceGuard g(h->device);
  const double* d_b = b;
  double* d_x = x;
  if (memspace == QRK_HOST) {
    int st = ensure_buffer(h, h->d_b, h->cap_b, (size_t)h->n_rows);
    if (st != QRK_STATUS_OK) return st;
    st = ensure_buffer(h, h->d_x, h->cap_x, (size_t)h->n_cols);
    if (st != QRK_STATUS_OK) return st;
    QRK_TRY_CUDA(h, cudaMemcpyAsync(h->d_b, b, h->n_rows * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    d_b = h->d_b; d_x = h->d_x;
  } else {
    QRK_REQUIRE(h, aligned16(b) && aligned16(x), "device b / x must be 16-byte aligned");
  }
  if (h->n_cols > h->sum_cols && !ang(h))
    QRK_TRY_CUDA(h, cudaMemsetAsync(d_x + h->sum_cols, 0, (h->n_cols - h->sum_cols) * sizeof(double), h->stream));
  int st = ang(h) ? angular_run(h, h->d_values, d_b, d_x, false) : (h->bvt ? banded_run(h, h->d_values, d_b, d_x) : run_factor(h, h->d_values, d_b, d_x));
  if (st != QRK_STATUS_OK) return st;
  h->factorized = true;
  if (h->pending) {            // multi-GPU block angular: x is produced by qrk_angular_merge
    h->pending_space = memspace;
    if (memspace == QRK_HOST) h->pending_x = x;
    return QRK_STATUS_OK;
  }
  if (memspace == QRK_HOST) {
    QRK_TRY_CUDA(h, cudaMemcpyAsync(x, h->d_x, h->n_cols * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
    return peer_exchange_status(h);
  }
  return QRK_STATUS_OK;
}

// Host buffers, thread-per-block kernels: the blocks are independent, so the upload, the fused kernel and the download
// of x are pipelined in chunks over three streams (H2D and D2H run concurrently on the two PCIe directions; the kernel
// time disappears behind the copies).  With pageable host memory the copies serialise and this degrades gracefully.
static int compute_solve_host_pipelined(qrk_solver* h, const double* values, const double* b, double* x) {
  DeviceGuard g(h->device);
  int st = ensure_own_values(h);
  if (st != QRK_STATUS_OK) return st;
  st = ensure_buffer(h, h->d_b, h->cap_b, (size_t)h->n_rows);
  if (st != QRK_STATUS_OK) return st;
  st = ensure_buffer(h, h->d_x, h->cap_x, (size_t)h->n_cols);
  if (st != QRK_STATUS_OK) return st;
  if (!h->analyzed) qrk_analyze_pattern(h, nullptr);
  const int r = h->ur, c = h->uc;
  const bool piv = h->desc.pivoting == QRK_PIVOT_COLPIV;
  const long long chunk = std::max<long long>(kSmallTPB, ((16LL << 20) / ((long long)r * c * 8)) / kSmallTPB * kSmallTPB);
  const int nchunks = (int)((h->nb + chunk - 1) / chunk);
  if (!h->s_in) {
    QRK_TRY_CUDA(h, cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    QRK_TRY_CUDA(h, cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  }
  while ((int)h->pipe_events.size() < 2 * nchunks + 1) {
    cudaEvent_t e;
    QRK_TRY_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->pipe_events.push_back(e);
  }
  cudaEvent_t ev0 = h->pipe_events[2 * nchunks];
  QRK_TRY_CUDA(h, cudaEventRecord(ev0, h->stream));            // earlier work on the handle's stream comes first
  QRK_TRY_CUDA(h, cudaStreamWaitEvent(h->s_in, ev0, 0));
  QRK_TRY_CUDA(h, cudaStreamWaitEvent(h->s_out, ev0, 0));
  // a failure in the middle must not leave copies in flight into the caller's buffers: drain all three streams first
  auto bail = [&](cudaError_t e, const char* what) {
    cudaStreamSynchronize(h->s_in); cudaStreamSynchronize(h->stream); cudaStreamSynchronize(h->s_out);
    (void)cudaGetLastError();
    h->err = std::string(what) + ": " + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? QRK_STATUS_ALLOC_FAILED : QRK_STATUS_CUDA_ERROR;
  };
#define QRK_PIPE(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return bail(e__, #expr); } while (0)
  for (int i = 0; i < nchunks; i++) {
    const long long b0 = i * chunk, cnt = std::min(chunk, h->nb - b0);
    cudaEvent_t ev_in = h->pipe_events[2 * i], ev_k = h->pipe_events[2 * i + 1];
    QRK_PIPE(cudaMemcpyAsync(h->d_values + b0 * r * c, values + b0 * r * c, (size_t)cnt * r * c * sizeof(double), cudaMemcpyHostToDevice, h->s_in));
    QRK_PIPE(cudaMemcpyAsync(h->d_b + b0 * r, b + b0 * r, (size_t)cnt * r * sizeof(double), cudaMemcpyHostToDevice, h->s_in));
    QRK_PIPE(cudaEventRecord(ev_in, h->s_in));
    QRK_PIPE(cudaStreamWaitEvent(h->stream, ev_in, 0));
    QRK_PIPE(launch_small_factor_dyn(r, c, piv, true, h->d_values + b0 * r * c, h->d_values + b0 * r * c, h->d_tau + b0 * c,
                                     h->d_perm + b0 * c, h->d_b + b0 * r, h->d_x + b0 * c, cnt, h->stream, b0));
    h->launches++;
    QRK_PIPE(cudaEventRecord(ev_k, h->stream));
    QRK_PIPE(cudaStreamWaitEvent(h->s_out, ev_k, 0));
    QRK_PIPE(cudaMemcpyAsync(x + b0 * c, h->d_x + b0 * c, (size_t)cnt * c * sizeof(double), cudaMemcpyDeviceToHost, h->s_out));
  }
#undef QRK_PIPE
  if (h->n_cols > h->sum_cols) std::fill(x + h->sum_cols, x + h->n_cols, 0.0);     // y.bottomRows(...).setZero() (:272)
  QRK_TRY_CUDA(h, cudaStreamSynchronize(h->s_out));
  QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
  h->has_blocks = true;
  h->factorized = true;
  return QRK_STATUS_OK;
}

int qrk_compute_solve(qrk_handle_t h, const double* values, const double* b, double* x, int memspace) {
  if (!h) return QRK_STATUS_INVALID_ARGUMENT;
  if (memspace == QRK_HOST && h->small_path && !h->bvt && !ang(h) && h->info != QRK_INFO_INVALID_INPUT && h->nb >= 4 * kSmallTPB) {
    QRK_REQUIRE(h, values && b && x, "values / b / x is null");
    return compute_solve_host_pipelined(h, values, b, x);
  }
  if (memspace == QRK_DEVICE) {
    QRK_REQUIRE(h, b && x, "b / x is null");
    QRK_REQUIRE(h, aligned16(b) && aligned16(x), "device b / x must be 16-byte aligned");
    return compute_from_device(h, values, b, x);
  }
  int st = qrk_set_blocks(h, values, memspace);
  if (st != QRK_STATUS_OK) return st;
  return qrk_factorize_solve(h, b, x, memspace);
}

int qrk_rows(qrk_handle_t h, int64_t* rows) { if (!h || !rows) return QRK_STATUS_INVALID_ARGUMENT; *rows = h->n_rows; return QRK_STATUS_OK; }
int qrk_cols(qrk_handle_t h, int64_t* cols) { if (!h || !cols) return QRK_STATUS_INVALID_ARGUMENT; *cols = h->n_cols; return QRK_STATUS_OK; }
int qrk_rank(qrk_handle_t h, int64_t* rank) {
  if (!h || !rank) return QRK_STATUS_INVALID_ARGUMENT;
  if (!h->factorized) return QRK_STATUS_NOT_FACTORIZED;
  *rank = h->sum_cols;   // rank += blockSolver.cols() (BlockDiagonalSparseQR.h:440)
  if (ang(h)) {          // + rightSolver.rank() (BlockAngularSparseQR.h:510)
    if (!h->root_done) return QRK_STATUS_NOT_FACTORIZED;
    DeviceGuard g(h->device);
    int r2 = 0;
    QRK_TRY_CUDA(h, cudaMemcpyAsync(&r2, h->d_root_i + h->m2, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
    *rank += r2;
    return peer_exchange_status(h);
  }
  return QRK_STATUS_OK;
}
int qrk_info(qrk_handle_t h, int32_t* info) { if (!h || !info) return QRK_STATUS_INVALID_ARGUMENT; *info = h->info; return QRK_STATUS_OK; }

int qrk_cols_permutation(qrk_handle_t h, int32_t* indices, int memspace) {
  if (!h || !indices) return QRK_STATUS_INVALID_ARGUMENT;
  if (!h->factorized || (ang(h) && !h->root_done)) return QRK_STATUS_NOT_FACTORIZED;
  DeviceGuard g(h->device);
  QRK_TRY_CUDA(h, cudaMemcpyAsync(indices, h->d_perm, h->n_cols * sizeof(int),
                                  memspace == QRK_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  if (memspace == QRK_HOST) QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
  return QRK_STATUS_OK;
}

int qrk_rows_permutation(qrk_handle_t h, int32_t* indices, int memspace) {
  if (!h || !indices) return QRK_STATUS_INVALID_ARGUMENT;
  if (!h->analyzed) return QRK_STATUS_NOT_FACTORIZED;
  DeviceGuard g(h->device);
  if (memspace == QRK_DEVICE) {
    QRK_TRY_CUDA(h, cudaMemcpyAsync(indices, h->h_rowperm.data(), h->n_rows * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
  } else {
    std::copy(h->h_rowperm.begin(), h->h_rowperm.end(), indices);
  }
  return QRK_STATUS_OK;
}

int qrk_packed_factors(qrk_handle_t h, double* packed, double* tau, int memspace) {
  if (!h) return QRK_STATUS_INVALID_ARGUMENT;
  if (!h->factorized) return QRK_STATUS_NOT_FACTORIZED;
  if (h->bgen) { h->err = "qrk_packed_factors: the general banded window chain keeps its reflectors per window, not in the block-COO layout"; return QRK_STATUS_UNSUPPORTED; }
  DeviceGuard g(h->device);
  const cudaMemcpyKind kind = memspace == QRK_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (packed) QRK_TRY_CUDA(h, cudaMemcpyAsync(packed, h->d_values, h->total_values * sizeof(double), kind, h->stream));
  if (tau) QRK_TRY_CUDA(h, cudaMemcpyAsync(tau, h->d_tau, h->n_cols * sizeof(double), kind, h->stream));
  if (memspace == QRK_HOST) QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
  return QRK_STATUS_OK;
}

// ---- matrixR / matrixQ export --------------------------------------------------------------------
int qrk_matrix_r_nnz(qrk_handle_t h, int64_t* nnz) {
  if (!h || !nnz) return QRK_STATUS_INVALID_ARGUMENT;
  if (!h->factorized) return QRK_STATUS_NOT_FACTORIZED;
  long long n = 0;
  if (h->uniform) n = h->nb * ((long long)h->uc * (h->uc + 1) / 2);
  else for (long long i = 0; i < h->nb; i++) n += (long long)h->h_cols[i] * (h->h_cols[i] + 1) / 2;
  if (ang(h)) n += h->sum_cols * (long long)h->m2 + (long long)h->m2 * (h->m2 + 1) / 2;   // border columns (makeR :296-305)
  if (h->bvt) n = banded_r_outer(h).back() + (ang(h) ? h->sum_cols * (long long)h->m2 + (long long)h->m2 * (h->m2 + 1) / 2 : 0);
  *nnz = n;
  return QRK_STATUS_OK;
}

int qrk_matrix_q_nnz(qrk_handle_t h, int64_t* nnz) {
  if (!h || !nnz) return QRK_STATUS_INVALID_ARGUMENT;
  if (!h->factorized) return QRK_STATUS_NOT_FACTORIZED;
  long long n = 0;
  if (h->uniform) n = h->nb * ((long long)h->ur * h->ur);
  else for (long long i = 0; i < h->nb; i++) n += (long long)h->h_rows[i] * h->h_rows[i];
  *nnz = n + (h->n_rows - h->sum_rows);
  return QRK_STATUS_OK;
}

static int export_sparse(qrk_handle_t h, bool want_q, int32_t* outer, int32_t* inner, double* values, int memspace) {
  if (!h || !outer || !inner || !values) return QRK_STATUS_INVALID_ARGUMENT;
  if (!h->factorized) return QRK_STATUS_NOT_FACTORIZED;
  if (ang(h) && !want_q && !h->root_done) return QRK_STATUS_NOT_FACTORIZED;
  int64_t nnz = 0;
  if (want_q) qrk_matrix_q_nnz(h, &nnz); else qrk_matrix_r_nnz(h, &nnz);
  QRK_REQUIRE(h, nnz <= INT32_MAX, "matrix has more than 2^31-1 stored entries (StorageIndex = int)");
  DeviceGuard g(h->device);
  const long long n_outer = (want_q ? h->n_rows : h->n_cols) + 1;
  int32_t *d_outer = outer, *d_inner = inner;
  double* d_vals = values;
  if (memspace == QRK_HOST) {
    d_outer = d_inner = nullptr; d_vals = nullptr;
    QRK_TRY_CUDA(h, cudaMalloc(&d_outer, n_outer * sizeof(int32_t)));
    if (cudaMalloc(&d_inner, std::max<long long>(nnz, 1) * sizeof(int32_t)) != cudaSuccess ||
        cudaMalloc(&d_vals, std::max<long long>(nnz, 1) * sizeof(double)) != cudaSuccess) {
      cudaFree(d_outer); if (d_inner) cudaFree(d_inner);
      h->err = "export: device allocation failed";
      return QRK_STATUS_ALLOC_FAILED;
    }
  }
  // per-block start of its entries in the compressed arrays (non-uniform: prefix sums on the host)
  long long* d_eoff = nullptr;
  if (!h->uniform && h->nb > 0) {
    std::vector<long long> eoff(h->nb);
    long long acc = 0;
    for (long long i = 0; i < h->nb; i++) {
      eoff[i] = acc;
      acc += want_q ? (long long)h->h_rows[i] * h->h_rows[i] : (long long)h->h_cols[i] * (h->h_cols[i] + 1) / 2;
    }
    cudaError_t ee = cudaMalloc(&d_eoff, h->nb * sizeof(long long));
    if (ee == cudaSuccess) ee = cudaMemcpyAsync(d_eoff, eoff.data(), h->nb * sizeof(long long), cudaMemcpyHostToDevice, h->stream);
    if (ee == cudaSuccess) ee = cudaStreamSynchronize(h->stream);
    if (ee != cudaSuccess) {                      // never launch the export kernels with a null / unfilled offset table
      (void)cudaGetLastError();
      if (d_eoff) cudaFree(d_eoff);
      if (memspace == QRK_HOST) { cudaFree(d_outer); cudaFree(d_inner); cudaFree(d_vals); }
      h->err = std::string("export: per-block offset table: ") + cudaGetErrorString(ee);
      return ee == cudaErrorMemoryAllocation ? QRK_STATUS_ALLOC_FAILED : QRK_STATUS_CUDA_ERROR;
    }
  }
  const BlockIndex bi = block_index(h);
  const int full_q = h->desc.q_format == QRK_FULL_Q ? 1 : 0;
  cudaError_t e;
  if (h->bvt) {
    if (want_q) { h->err = "banded matrixQ() as an explicit sparse matrix is not provided"; e = cudaErrorNotSupported; }
    else {
      std::vector<int> ho, hi;
      banded_r_pattern(h, ho, &hi);
      cudaMemcpyAsync(d_outer, ho.data(), ho.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream);
      cudaMemcpyAsync(d_inner, hi.data(), hi.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream);
      if (h->bgen) e = banded_generic_export_r(h->g_args, h->d_gcol0, d_outer, d_inner, d_vals, h->stream);
      else {
        export_banded_r_kernel<<<148 * 4, 256, 0, h->stream>>>(h->d_rband, d_outer, d_inner, d_vals, h->sum_cols, h->nb, h->uc, h->b_step);
        e = cudaGetLastError();
      }
      if (ang(h) && e == cudaSuccess) {     // R = [R1, Atop P2; 0, R2] (makeR, BlockAngularSparseQR.h:285-308) with a banded R1
        export_angular_border_kernel<<<148, 256, 0, h->stream>>>(h->d_wx, h->w_ld, h->d_root, h->d_root_i, h->sum_cols, h->m2, (long long)ho.back(),
                                                                 d_outer, d_inner, d_vals);
        e = cudaGetLastError();
        h->launches++;
      }
      cudaStreamSynchronize(h->stream);     // ho must outlive the copy
    }
  } else if (want_q) e = launch_export_q(bi, d_eoff, h->nb, h->d_values, h->d_tau, h->n_rows, ang(h) ? h->sum_cols : h->n_cols, h->sum_rows,
                                  nnz - (h->n_rows - h->sum_rows), full_q, h->max_r, d_outer, d_inner, d_vals, h->stream);
  else if (ang(h)) {
    const long long nnz_r1 = nnz - (h->sum_cols * (long long)h->m2 + (long long)h->m2 * (h->m2 + 1) / 2);
    e = launch_export_r(bi, d_eoff, h->nb, h->d_values, h->sum_cols, h->sum_cols, nnz_r1, 1, d_outer, d_inner, d_vals, h->stream);
    export_angular_border_kernel<<<148, 256, 0, h->stream>>>(h->wide ? h->d_wx : h->d_atop, h->wide ? h->w_ld : h->sum_cols, h->d_root,
                                                             h->d_root_i, h->sum_cols, h->m2, nnz_r1, d_outer,
                                                             d_inner, d_vals);
    if (e == cudaSuccess) e = cudaGetLastError();
    h->launches++;
  } else e = launch_export_r(bi, d_eoff, h->nb, h->d_values, h->n_cols, h->sum_cols, nnz, full_q, d_outer, d_inner, d_vals, h->stream);
  h->launches += 2;
  int st = QRK_STATUS_OK;
  if (e != cudaSuccess) { h->err = std::string("export kernel: ") + cudaGetErrorString(e); st = QRK_STATUS_CUDA_ERROR; }
  if (st == QRK_STATUS_OK && memspace == QRK_HOST) {
    cudaMemcpyAsync(outer, d_outer, n_outer * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(inner, d_inner, nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(values, d_vals, nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  }
  e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess && st == QRK_STATUS_OK) { h->err = std::string("export: ") + cudaGetErrorString(e); st = QRK_STATUS_CUDA_ERROR; }
  if (memspace == QRK_HOST) { cudaFree(d_outer); cudaFree(d_inner); cudaFree(d_vals); }
  if (d_eoff) cudaFree(d_eoff);
  return st;
}

int qrk_matrix_r(qrk_handle_t h, int32_t* outer, int32_t* inner, double* values, int memspace) {
  return export_sparse(h, false, outer, inner, values, memspace);
}
int qrk_matrix_q(qrk_handle_t h, int32_t* outer, int32_t* inner, double* values, int memspace) {
  return export_sparse(h, true, outer, inner, values, memspace);
}

// ---- Q^T B, Q B, solve ----------------------------------------------------------------------------
constexpr int OP_APPLY_QT_THIN = 10, OP_APPLY_Q_THIN = 11;   // Q1^T B (n_cols rows out) / Q1 Y (n_cols rows in): the thin factor

static int op_entry(qrk_handle_t h, int op, const double* B, int64_t ldb, double* X, int64_t ldx, int32_t nrhs, int memspace) {
  if (!h) return QRK_STATUS_INVALID_ARGUMENT;
  if (!h->factorized) return QRK_STATUS_NOT_FACTORIZED;
  QRK_REQUIRE(h, B && X && nrhs >= 0, "B / X is null or nrhs < 0");
  const bool thin = op == OP_APPLY_QT_THIN || op == OP_APPLY_Q_THIN;
  const long long in_rows = (op == OP_APPLY_Q_THIN) ? h->n_cols : h->n_rows;
  const long long out_rows = (op == OP_SOLVE || op == OP_APPLY_QT_THIN) ? h->n_cols : h->n_rows;
  QRK_REQUIRE(h, ldb >= in_rows && ldx >= out_rows, "leading dimension smaller than the number of rows");
  const bool banded_q1 = h->bvt != nullptr;               // banded solver, or block angular with a banded left solver
  if (banded_q1 && !h->bgen && (op == OP_APPLY_QT || op == OP_APPLY_Q)) {
    // The two-phase banded factorisation represents Q as an isometry into an EXTENDED complement (overlap rows of every
    // group enter as virtual zero rows, banded.cuh): there is no n x n orthogonal matrix to multiply with.  The thin factor
    // Q1 = A R^-1 (what solve() and the LM caller use) is exact: qrk_apply_qt_thin / qrk_apply_q_thin.
    h->err = "matrixQ() as an n x n operator is not provided for a banded factor; use qrk_apply_qt_thin / qrk_apply_q_thin (Q1 = A R^-1)";
    return QRK_STATUS_UNSUPPORTED;
  }
  if (thin && !banded_q1 && h->desc.q_format != QRK_FULL_Q) {
    h->err = "the thin-factor products need the FullQ index layout (thin part first)";
    return QRK_STATUS_UNSUPPORTED;
  }
  if (op == OP_APPLY_Q_THIN && ang(h) && h->left_banded) {
    h->err = "qrk_apply_q_thin is not provided for a block-angular handle with a banded left solver";
    return QRK_STATUS_UNSUPPORTED;
  }
  DeviceGuard g(h->device);
  const double* d_B = B;
  double* d_X = X;
  long long dldb = ldb, dldx = ldx;
  if (memspace == QRK_HOST) {
    dldb = (in_rows + 1) & ~1LL; dldx = (out_rows + 1) & ~1LL;
    int st = ensure_buffer(h, h->d_b, h->cap_b, (size_t)dldb * nrhs);
    if (st != QRK_STATUS_OK) return st;
    st = ensure_buffer(h, h->d_x, h->cap_x, (size_t)dldx * nrhs);
    if (st != QRK_STATUS_OK) return st;
    QRK_TRY_CUDA(h, cudaMemcpy2DAsync(h->d_b, dldb * sizeof(double), B, ldb * sizeof(double), in_rows * sizeof(double), nrhs,
                                      cudaMemcpyHostToDevice, h->stream));
    d_B = h->d_b; d_X = h->d_x;
  }
  int st = QRK_STATUS_OK;
  if (thin && !banded_q1) {
    // thin products through the full operator on a scratch copy: Q1^T B = (Q^T B)[0:n_cols], Q1 Y = Q [Y; 0]
    const long long ldt = (h->n_rows + 1) & ~1LL;
    st = ensure_buffer(h, h->d_qthin, h->cap_qthin, (size_t)ldt * nrhs);
    if (st != QRK_STATUS_OK) return st;
    const int full_op = (op == OP_APPLY_QT_THIN) ? OP_APPLY_QT : OP_APPLY_Q;
    const double* src = d_B;
    long long lds = dldb;
    double* dst = d_X;
    long long ldd = dldx;
    if (op == OP_APPLY_QT_THIN) { dst = h->d_qthin; ldd = ldt; }
    else {
      QRK_TRY_CUDA(h, cudaMemsetAsync(h->d_qthin, 0, (size_t)ldt * nrhs * sizeof(double), h->stream));
      QRK_TRY_CUDA(h, cudaMemcpy2DAsync(h->d_qthin, ldt * sizeof(double), d_B, dldb * sizeof(double), h->n_cols * sizeof(double), nrhs,
                                        cudaMemcpyDeviceToDevice, h->stream));
      src = h->d_qthin; lds = ldt;
    }
    if (h->n_rows > h->sum_rows) { copy_tail_kernel<<<64, 256, 0, h->stream>>>(src, lds, dst, ldd, nrhs, h->sum_rows, h->n_rows); h->launches++; }
    st = ang(h) ? angular_apply(h, full_op, src, lds, dst, ldd, nrhs) : run_op(h, full_op, src, lds, dst, ldd, nrhs);
    if (st != QRK_STATUS_OK) return st;
    if (op == OP_APPLY_QT_THIN)
      QRK_TRY_CUDA(h, cudaMemcpy2DAsync(d_X, dldx * sizeof(double), h->d_qthin, ldt * sizeof(double), h->n_cols * sizeof(double), nrhs,
                                        cudaMemcpyDeviceToDevice, h->stream));
  } else if (thin && ang(h)) {
    // banded left solver: [Q1thin^T b ; (Q2^T of the extended complement)[0:m2]] — what _solve_impl forms before the triangular solves
    const long long n = h->w_ld, m1 = h->sum_cols;
    const int M = h->m2;
    double* rhs_col = h->d_wx + (long long)M * n;
    for (int j = 0; j < nrhs && st == QRK_STATUS_OK; j++) {
      st = banded_left_apply_qt(h, d_B + j * dldb, h->n_rows, rhs_col, 1);
      if (st == QRK_STATUS_OK) st = apply_q2(h, rhs_col + m1, true);
      if (st == QRK_STATUS_OK)
        QRK_TRY_CUDA(h, cudaMemcpyAsync(d_X + j * dldx, rhs_col, (size_t)(m1 + M) * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
  } else if (op == OP_SOLVE) {
    if (h->n_cols > h->sum_cols && !ang(h))   // y.bottomRows(...).setZero() (:272)
      QRK_TRY_CUDA(h, cudaMemset2DAsync(d_X + h->sum_cols, dldx * sizeof(double), 0, (h->n_cols - h->sum_cols) * sizeof(double),
                                        nrhs, h->stream));
  } else if (!thin && h->n_rows > h->sum_rows && !h->bvt) {
    copy_tail_kernel<<<64, 256, 0, h->stream>>>(d_B, dldb, d_X, dldx, nrhs, h->sum_rows, h->n_rows);
    h->launches++;
  }
  if (thin && !(banded_q1 && !ang(h))) {
    // done above
  } else if (h->bvt && !ang(h)) {
    // Q1^T b by the window sweep over the stored reflectors (BandedBlockedSparseQR.h:655-670 applies the YTY blocks in
    // the same order) and the chase reflectors; solve: + the banded back substitution (:299-304).
    // Q1 y: the same reflectors in reverse order on [y; 0].
    for (int j = 0; j < nrhs && st == QRK_STATUS_OK; j++) {
      BandedArgs a = banded_args(h);
      cudaError_t e;
      if (h->bgen && (op == OP_APPLY_QT || op == OP_APPLY_Q)) {
        // the general window chain has an exact n x n Q: [thin part ; complement in window order] (banded_generic.cuh)
        if (op == OP_APPLY_QT) { a.b = d_B + j * dldb; a.y = d_X + j * dldx; a.comp = d_X + j * dldx + h->sum_cols; e = h->bvt->apply_qt(a, h->stream); }
        else { a.y = const_cast<double*>(d_B + j * dldb); a.comp = const_cast<double*>(d_B + j * dldb + h->sum_cols); a.x = d_X + j * dldx; e = h->bvt->apply_q(a, h->stream); }
        h->launches++;
      } else if (op == OP_APPLY_Q_THIN) {
        a.y = const_cast<double*>(d_B + j * dldb);   // read only: the thin part of the input
        a.x = d_X + j * dldx;
        e = h->bvt->apply_q(a, h->stream);
        h->launches += banded_launches_per_call();
      } else {
        a.b = d_B + j * dldb;
        if (op == OP_APPLY_QT_THIN) a.y = d_X + j * dldx;
        e = h->bvt->apply_qt(a, h->stream);
        h->launches += banded_launches_per_call();
        if (e == cudaSuccess && op == OP_SOLVE) { a.x = d_X + j * dldx; e = h->bvt->backsolve(a, h->stream); h->launches++; }
      }
      if (e != cudaSuccess) { h->err = std::string("banded op: ") + cudaGetErrorString(e); st = QRK_STATUS_CUDA_ERROR; }
    }
  } else if (ang(h) && op == OP_SOLVE) {
    QRK_REQUIRE(h, h->world == 1 || nrhs == 1, "multi-GPU block-angular solve takes one right-hand side per call");
    for (int j = 0; j < nrhs && st == QRK_STATUS_OK; j++) st = angular_solve_stored(h, d_B + j * dldb, d_X + j * dldx);
    if (st == QRK_STATUS_OK && h->pending) {
      h->pending_space = memspace;
      if (memspace == QRK_HOST) h->pending_x = X;
      return QRK_STATUS_OK;
    }
  } else if (ang(h)) {
    st = angular_apply(h, op, d_B, dldb, d_X, dldx, nrhs);
  } else {
    st = run_op(h, op, d_B, dldb, d_X, dldx, nrhs);
  }
  if (st != QRK_STATUS_OK) return st;
  if (memspace == QRK_HOST) {
    QRK_TRY_CUDA(h, cudaMemcpy2DAsync(X, ldx * sizeof(double), h->d_x, dldx * sizeof(double), out_rows * sizeof(double), nrhs,
                                      cudaMemcpyDeviceToHost, h->stream));
    QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
    if (int pst = peer_exchange_status(h)) return pst;
  }
  h->info = QRK_INFO_SUCCESS;   // m_info = Success (:278)
  return QRK_STATUS_OK;
}

int qrk_apply_qt_thin(qrk_handle_t h, const double* B, int64_t ldb, double* Y, int64_t ldy, int32_t nrhs, int memspace) {
  return op_entry(h, OP_APPLY_QT_THIN, B, ldb, Y, ldy, nrhs, memspace);
}
int qrk_apply_q_thin(qrk_handle_t h, const double* Y, int64_t ldy, double* X, int64_t ldx, int32_t nrhs, int memspace) {
  return op_entry(h, OP_APPLY_Q_THIN, Y, ldy, X, ldx, nrhs, memspace);
}
int qrk_apply_qt(qrk_handle_t h, const double* B, int64_t ldb, double* Y, int64_t ldy, int32_t nrhs, int memspace) {
  return op_entry(h, OP_APPLY_QT, B, ldb, Y, ldy, nrhs, memspace);
}
int qrk_apply_q(qrk_handle_t h, const double* B, int64_t ldb, double* Y, int64_t ldy, int32_t nrhs, int memspace) {
  return op_entry(h, OP_APPLY_Q, B, ldb, Y, ldy, nrhs, memspace);
}
int qrk_solve(qrk_handle_t h, const double* B, int64_t ldb, double* X, int64_t ldx, int32_t nrhs, int memspace) {
  return op_entry(h, OP_SOLVE, B, ldb, X, ldx, nrhs, memspace);
}

// ---- block angular entry points --------------------------------------------------------------------
int qrk_set_border(qrk_handle_t h, const double* J2, int64_t ld, int memspace) {
  if (!h) return QRK_STATUS_INVALID_ARGUMENT;
  QRK_REQUIRE(h, ang(h), "qrk_set_border: the handle is not of kind QRK_BLOCK_ANGULAR");
  QRK_REQUIRE(h, J2 && ld >= h->n_rows, "border is null or its leading dimension is smaller than the number of rows");
  DeviceGuard g(h->device);
  if (memspace == QRK_DEVICE) {
    QRK_REQUIRE(h, (reinterpret_cast<uintptr_t>(J2) & 7) == 0, "border must be 8-byte aligned");
    QRK_REQUIRE(h, (h->ur % 2) || (aligned16(J2) && ld % 2 == 0),
                "even block_rows: the device border must be 16-byte aligned with an even leading dimension");
    h->d_border = J2;
    h->ld_border = ld;
  } else {
    const size_t need = (size_t)h->n_rows * h->m2;
    int st = ensure_buffer(h, h->d_border_own, h->cap_border, need);
    if (st != QRK_STATUS_OK) return st;
    QRK_TRY_CUDA(h, cudaMemcpy2DAsync(h->d_border_own, h->n_rows * sizeof(double), J2, ld * sizeof(double),
                                      h->n_rows * sizeof(double), h->m2, cudaMemcpyHostToDevice, h->stream));
    h->d_border = h->d_border_own;
    h->ld_border = h->n_rows;
  }
  return QRK_STATUS_OK;
}

int qrk_angular_set_world(qrk_handle_t h, int32_t world_size) {
  if (!h || !h->avt || world_size < 1) return QRK_STATUS_INVALID_ARGUMENT;
  h->world = world_size;
  h->xchg_rank = -1;                 // a new world needs a new qrk_angular_p2p_attach
  return QRK_STATUS_OK;
}

int qrk_angular_triangle_size(qrk_handle_t h, int64_t* doubles) {
  if (!h || !h->avt || !doubles) return QRK_STATUS_INVALID_ARGUMENT;
  *doubles = h->avt->tri_doubles;
  return QRK_STATUS_OK;
}

int qrk_angular_local_triangle(qrk_handle_t h, double* tri, int memspace) {
  if (!h || !h->avt || !tri) return QRK_STATUS_INVALID_ARGUMENT;
  if (!h->factorized) return QRK_STATUS_NOT_FACTORIZED;
  QRK_REQUIRE(h, h->world > 1, "the per-GPU triangle is only kept when qrk_angular_set_world(world > 1) was called");
  DeviceGuard g(h->device);
  QRK_TRY_CUDA(h, cudaMemcpyAsync(tri, h->d_tri, h->avt->tri_doubles * sizeof(double),
                                  memspace == QRK_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  if (memspace == QRK_HOST) QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
  return QRK_STATUS_OK;
}

int qrk_angular_merge(qrk_handle_t h, const double* tris, int32_t count, int memspace) {
  if (!h || !h->avt || !tris || count < 1) return QRK_STATUS_INVALID_ARGUMENT;
  if (!h->factorized || !h->pending) return QRK_STATUS_NOT_FACTORIZED;
  DeviceGuard g(h->device);
  const size_t n = (size_t)count * h->avt->tri_doubles;
  const double* d_tris = tris;
  double* tmp = nullptr;
  if (memspace == QRK_HOST) {
    QRK_TRY_CUDA(h, cudaMalloc(&tmp, n * sizeof(double)));
    cudaMemcpyAsync(tmp, tris, n * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    d_tris = tmp;
  }
  AngularArgs a = angular_args(h);
  a.tris = d_tris; a.tri_count = count; a.tris_ld = 0; a.root_mode = 1; a.keep_rhs_only = h->pending_keep_rhs_only;
  int st = QRK_STATUS_OK;
  cudaError_t e = h->avt->root(a, h->stream);
  h->launches++;
  h->root_done = true;
  if (e == cudaSuccess && h->pending_x) {
    double* d_x = (h->pending_space == QRK_HOST) ? h->d_x : h->pending_x;
    a.x = d_x;
    e = h->avt->backsolve(a, h->stream);
    h->launches++;
    if (e == cudaSuccess && h->pending_space == QRK_HOST)
      e = cudaMemcpyAsync(h->pending_x, h->d_x, h->n_cols * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  }
  if (e == cudaSuccess && (memspace == QRK_HOST || h->pending_space == QRK_HOST)) e = cudaStreamSynchronize(h->stream);
  if (tmp) { cudaStreamSynchronize(h->stream); cudaFree(tmp); }
  if (e != cudaSuccess) { h->err = std::string("qrk_angular_merge: ") + cudaGetErrorString(e); st = QRK_STATUS_CUDA_ERROR; }
  h->pending = false;
  h->pending_x = nullptr;
  return st;
}

int qrk_angular_xchg_buffer(qrk_handle_t h, void** device_ptr, int64_t* bytes) {
  if (!h || !h->avt || !device_ptr || !bytes) return QRK_STATUS_INVALID_ARGUMENT;
  QRK_REQUIRE(h, h->world > 1 && !h->wide, "call qrk_angular_set_world(world > 1) first (fused TSQR path only)");
  DeviceGuard g(h->device);
  const size_t need = (size_t)2 * h->world * h->avt->tri_doubles * sizeof(double) + (size_t)2 * h->world * sizeof(unsigned long long);
  if (h->d_xchg && h->xchg_bytes != need) { cudaFree(h->d_xchg); h->d_xchg = nullptr; }
  if (!h->d_xchg) {
    QRK_TRY_CUDA(h, cudaMalloc(&h->d_xchg, need));          // a plain cudaMalloc allocation: exportable with cudaIpcGetMemHandle
    QRK_TRY_CUDA(h, cudaMemset(h->d_xchg, 0, need));        // flags = 0 < every step number
    h->xchg_bytes = need;
  }
  *device_ptr = h->d_xchg;
  *bytes = (int64_t)need;
  return QRK_STATUS_OK;
}

int qrk_angular_p2p_attach(qrk_handle_t h, void* const* peer_buffers, int32_t world_size, int32_t rank) {
  if (!h || !h->avt || !peer_buffers) return QRK_STATUS_INVALID_ARGUMENT;
  QRK_REQUIRE(h, world_size == h->world && world_size > 1 && rank >= 0 && rank < world_size, "world / rank do not match qrk_angular_set_world");
  QRK_REQUIRE(h, h->d_xchg && peer_buffers[rank] == (void*)h->d_xchg, "peer_buffers[rank] must be this handle's qrk_angular_xchg_buffer");
  for (int g = 0; g < world_size; g++) QRK_REQUIRE(h, peer_buffers[g] != nullptr, "null peer buffer");
  DeviceGuard g(h->device);
  if (!h->d_xchg_peers) QRK_TRY_CUDA(h, cudaMalloc(&h->d_xchg_peers, (size_t)world_size * sizeof(double*)));
  if (!h->d_xchg_err) { QRK_TRY_CUDA(h, cudaMalloc(&h->d_xchg_err, sizeof(int))); QRK_TRY_CUDA(h, cudaMemset(h->d_xchg_err, 0, sizeof(int))); }
  QRK_TRY_CUDA(h, cudaMemcpy(h->d_xchg_peers, peer_buffers, (size_t)world_size * sizeof(double*), cudaMemcpyHostToDevice));
  QRK_TRY_CUDA(h, h->avt->preload(h->ur, h->uc, h->desc.pivoting == QRK_PIVOT_COLPIV));   // no lazy load behind a spinning kernel
  // ... and no lazy ALLOCATION either: cudaMalloc / cudaFree wait for every running kernel of the device, and a root kernel
  // that spins for a peer never finishes while that peer's thread sits in cudaMalloc (two ranks on one device: a stall
  // until the timeout).  Everything the host- and device-memspace calls of this handle allocate on first use is set up here.
  {
    int st = h->d_values ? QRK_STATUS_OK : ensure_own_values(h);     // (adopted caller storage stays)
    if (st == QRK_STATUS_OK) st = ensure_buffer(h, h->d_b, h->cap_b, (size_t)h->n_rows);
    if (st == QRK_STATUS_OK) st = ensure_buffer(h, h->d_x, h->cap_x, (size_t)h->n_cols);
    if (st == QRK_STATUS_OK) st = ensure_buffer(h, h->d_border_own, h->cap_border, (size_t)h->n_rows * h->m2);
    if (st != QRK_STATUS_OK) return st;
    if (!h->d_abot) QRK_TRY_CUDA(h, cudaMalloc(&h->d_abot, std::max<long long>(1, (h->n_rows - h->sum_cols) * (long long)(h->m2 + 1)) * sizeof(double)));
  }
  QRK_TRY_CUDA(h, cudaMemset(h->d_xchg_err, 0, sizeof(int)));
  if (!h->d_xchg_seq) QRK_TRY_CUDA(h, cudaMalloc(&h->d_xchg_seq, sizeof(unsigned long long)));
  QRK_TRY_CUDA(h, cudaMemset(h->d_xchg_seq, 0, sizeof(unsigned long long)));
  QRK_TRY_CUDA(h, cudaMemset(h->d_xchg, 0, h->xchg_bytes));       // flags back to 0: attach restarts the step count on every rank
  h->xchg_rank = rank;
  return QRK_STATUS_OK;
}

int qrk_angular_p2p_set_timeout(qrk_handle_t h, double seconds) {
  if (!h || !h->avt || !(seconds > 0.0)) return QRK_STATUS_INVALID_ARGUMENT;
  h->xchg_timeout_ns = (unsigned long long)(seconds * 1e9);
  return QRK_STATUS_OK;
}

int qrk_angular_p2p_status(qrk_handle_t h, int32_t* timed_out) {
  if (!h || !timed_out) return QRK_STATUS_INVALID_ARGUMENT;
  *timed_out = 0;
  if (!h->d_xchg_err) return QRK_STATUS_OK;
  DeviceGuard g(h->device);
  QRK_TRY_CUDA(h, cudaStreamSynchronize(h->stream));
  QRK_TRY_CUDA(h, cudaMemcpy(timed_out, h->d_xchg_err, sizeof(int), cudaMemcpyDeviceToHost));
  return QRK_STATUS_OK;
}

int qrk_ipc_export(const void* device_ptr, void* handle64) {
  if (!device_ptr || !handle64) return QRK_STATUS_INVALID_ARGUMENT;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C ABI passes IPC handles as 64 bytes");
  cudaIpcMemHandle_t hd;
  if (cudaIpcGetMemHandle(&hd, const_cast<void*>(device_ptr)) != cudaSuccess) { (void)cudaGetLastError(); return QRK_STATUS_CUDA_ERROR; }
  std::memcpy(handle64, &hd, 64);
  return QRK_STATUS_OK;
}

int qrk_ipc_import(const void* handle64, void** device_ptr) {
  if (!device_ptr || !handle64) return QRK_STATUS_INVALID_ARGUMENT;
  cudaIpcMemHandle_t hd;
  std::memcpy(&hd, handle64, 64);
  if (cudaIpcOpenMemHandle(device_ptr, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); return QRK_STATUS_CUDA_ERROR; }
  return QRK_STATUS_OK;
}

int qrk_ipc_close(void* device_ptr) {
  if (!device_ptr) return QRK_STATUS_INVALID_ARGUMENT;
  return cudaIpcCloseMemHandle(device_ptr) == cudaSuccess ? QRK_STATUS_OK : QRK_STATUS_CUDA_ERROR;
}

int qrk_enable_peer_access(int32_t device, int32_t peer) {
  int ndev = 0;
  qrk_device_count(&ndev);
  if (ndev <= 0) return QRK_STATUS_NO_DEVICE;
  if (device < 0 || device >= ndev || peer < 0 || peer >= ndev) return QRK_STATUS_INVALID_ARGUMENT;
  if (device == peer) return QRK_STATUS_OK;
  DeviceGuard g(device);
  int can = 0;
  if (cudaDeviceCanAccessPeer(&can, device, peer) != cudaSuccess || !can) { (void)cudaGetLastError(); return QRK_STATUS_UNSUPPORTED; }
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return QRK_STATUS_CUDA_ERROR; }
  (void)cudaGetLastError();
  return QRK_STATUS_OK;
}

int qrk_launch_count(qrk_handle_t h, int64_t* launches) {
  if (!h || !launches) return QRK_STATUS_INVALID_ARGUMENT;
  *launches = h->launches;
  return QRK_STATUS_OK;
}

int qrk_synth_fill(double* device_out, uint64_t seed, int64_t block0, int64_t nb, int32_t r, int32_t c, double lo, double hi,
                   void* cuda_stream) {
  if (!device_out || nb < 0 || r <= 0 || c < 0) return QRK_STATUS_INVALID_ARGUMENT;
  int ndev = 0;
  qrk_device_count(&ndev);
  if (ndev <= 0) return QRK_STATUS_NO_DEVICE;
  if (nb == 0) return QRK_STATUS_OK;
  synth_fill_kernel<<<148 * 8, 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(device_out, seed, block0, nb, r, c, lo, hi);
  return cudaGetLastError() == cudaSuccess ? QRK_STATUS_OK : QRK_STATUS_CUDA_ERROR;
}

int qrk_ellipse_points(double* px, double* py, int64_t n, double a, double b, double x0, double y0, double r, void* cuda_stream) {
  if (!px || !py || n <= 0) return QRK_STATUS_INVALID_ARGUMENT;
  int ndev = 0;
  qrk_device_count(&ndev);
  if (ndev <= 0) return QRK_STATUS_NO_DEVICE;
  ellipse_points_kernel<<<148 * 4, 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(px, py, n, a, b, x0, y0, r);
  return cudaGetLastError() == cudaSuccess ? QRK_STATUS_OK : QRK_STATUS_CUDA_ERROR;
}

int qrk_ellipse_assemble(const double* px, const double* py, const double* params, int64_t n, double* J1, double* J2, double* rhs,
                         double* cost, void* cuda_stream) {
  if (!px || !py || !params || !J1 || !J2 || !rhs || n <= 0) return QRK_STATUS_INVALID_ARGUMENT;
  if (!aligned16(J1) || !aligned16(J2) || !aligned16(rhs)) return QRK_STATUS_INVALID_ARGUMENT;
  int ndev = 0;
  qrk_device_count(&ndev);
  if (ndev <= 0) return QRK_STATUS_NO_DEVICE;
  const unsigned grid = (unsigned)std::min<long long>((n + 255) / 256, 148LL * 8);
  ellipse_assemble_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(px, py, params, n, J1, J2, rhs, cost);
  return cudaGetLastError() == cudaSuccess ? QRK_STATUS_OK : QRK_STATUS_CUDA_ERROR;
}

}  // extern "C"
