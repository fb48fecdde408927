// angular_dispatch.hpp — run-time dispatch to the block-angular kernels instantiated per border width M2.
// angular_inst.cu is compiled once per M2 (-DQRK_M2=k) so that the instantiations build in parallel.
#pragma once
#include <cuda_runtime.h>

namespace qrk {

struct AngularArgs {
  // left blocks
  int r = 0, c = 0;
  bool piv = false;
  long long nb = 0;
  const double* A_in = nullptr;
  double* packed = nullptr;
  double* tau = nullptr;
  int* perm = nullptr;
  // border
  const double* J2 = nullptr;
  long long ldj = 0;
  const double* b = nullptr;
  double* atop = nullptr;
  double* y1 = nullptr;
  double* abot = nullptr;
  // TSQR
  double* partials = nullptr;
  int grid = 0;            // CTAs of the factor / rhs kernel = number of partial triangles
  const double* tris = nullptr;
  int tri_count = 0;
  int tris_ld = 0;         // > 0: component-major (entry i of triangle q at i * tris_ld + q); 0: triangle-major
  int root_mode = 1;
  int keep_rhs_only = 0;
  double* out_tri = nullptr;
  double* root = nullptr;
  int* root_i = nullptr;
  int* perm_tail = nullptr;  // colsPermutation()[m1 .. m1+m2)
  int m1 = 0;
  // solution
  double* x = nullptr;
  // fused peer exchange (root_mode 2)
  double* const* xchg_peers = nullptr;
  int xchg_world = 0, xchg_rank = 0;
  unsigned long long* xchg_seq = nullptr;
  int* xchg_err = nullptr;
  unsigned long long xchg_timeout_ns = 0;   // 0: the kernel's default (10 s)
};

struct AngularVTable {
  int m2;
  int tri_doubles;                                                 // Tri<M2>::N
  bool (*shape_ok)(int r, int c);
  cudaError_t (*max_grid)(int r, int c, bool piv, int* grid);       // resident CTAs of the factor kernel on this device
  int (*tile_blocks)(int r, int c);                                // diagonal blocks per tile of the factor kernel
  cudaError_t (*factor)(const AngularArgs&, cudaStream_t);
  cudaError_t (*rhs)(const AngularArgs&, cudaStream_t);
  cudaError_t (*root)(const AngularArgs&, cudaStream_t);
  cudaError_t (*backsolve)(const AngularArgs&, cudaStream_t);
  cudaError_t (*preload)(int r, int c, bool piv);                   // force the (lazily loaded) kernels of this shape into the context
};

constexpr int kAngularMaxM2 = 8;
const AngularVTable* angular_vtable(int m2);   // nullptr when m2 is not instantiated

}  // namespace qrk
