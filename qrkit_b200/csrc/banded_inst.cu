// banded_inst.cu — instantiations of the block-banded kernels.
#include "banded.cuh"
#include "banded_generic.cuh"
#include "banded_dispatch.hpp"

namespace qrk {
namespace {

// (block_rows, block_cols, overlap): BASELINE config 4 (16, 24, 16); the reference test pattern 7x4 with overlap 2
// (test/test-qrkit.cpp:63-96) and the non-overlapping 7x2 pattern (:101-131); a few more for coverage.
#define QRK_BANDED_SHAPES(X) X(16, 24, 16) X(7, 4, 2) X(7, 2, 0) X(8, 8, 4) X(12, 8, 4) X(4, 6, 4)

inline unsigned groups_of(const BandedArgs& a) { return (unsigned)((a.nb + a.group - 1) / a.group); }

// Without overlap (OV = 0) consecutive slabs share no column: the groups are independent, the group triangles ARE the rows
// of R (the two buffers have the same layout) and there is nothing to chase — phase 2 is skipped in all three operations.
template <int BR, int BC, int OV>
cudaError_t factor_t(const BandedArgs& a, cudaStream_t s) {
  if constexpr (OV == 0) {
    banded_factor_kernel<BR, BC, OV><<<groups_of(a), 32, 0, s>>>(a.A_in, a.packed, a.tau, a.rband, a.b, a.y, a.nb, a.last_cols, a.group);
    return cudaGetLastError();
  }
  banded_factor_kernel<BR, BC, OV><<<groups_of(a), 32, 0, s>>>(a.A_in, a.packed, a.tau, a.gband, a.b, a.gy, a.nb, a.last_cols, a.group);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  banded_chase_kernel<BC, OV><<<1, 32, 0, s>>>(a.gband, a.gy, a.rband, a.y, a.cvec, a.ctau, a.nb, a.last_cols, a.group, a.b != nullptr);
  return cudaGetLastError();
}
template <int BR, int BC, int OV>
cudaError_t apply_qt_t(const BandedArgs& a, cudaStream_t s) {
  constexpr int S = BC - OV;
  const long long ldgy = (long long)groups_of(a) * ((long long)(a.group - 1) * S + BC);
  if constexpr (OV == 0) {                      // the pivot-row values of the groups are the thin part itself
    banded_apply_qt_kernel<BR, BC, OV><<<dim3(groups_of(a), (unsigned)a.ncols), 32, 0, s>>>(a.packed, a.tau, a.b, a.y, a.nb, a.last_cols, a.group, a.comp,
                                                                                     a.ldb, a.ldy, a.ldcomp);
    return cudaGetLastError();
  }
  banded_apply_qt_kernel<BR, BC, OV><<<dim3(groups_of(a), (unsigned)a.ncols), 32, 0, s>>>(a.packed, a.tau, a.b, a.gy, a.nb, a.last_cols, a.group, a.comp,
                                                                                   a.ldb, ldgy, a.ldcomp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  double* ucomp = a.comp ? a.comp + a.nb * (long long)(OV + BR - BC) : nullptr;
  if (a.ncols == 1)
    banded_chase_apply_kernel<BC, OV, false><<<1, 32, 0, s>>>(a.cvec, a.ctau, a.gy, a.y, a.nb, a.last_cols, a.group, ucomp);
  else
    banded_chase_apply_multi_kernel<BC, OV><<<(unsigned)((a.ncols + 31) / 32), 32, 0, s>>>(a.cvec, a.ctau, a.gy, ldgy, a.y, a.ldy, ucomp, a.ldcomp, a.ncols,
                                                                                     a.nb, a.last_cols, a.group);
  return cudaGetLastError();
}
template <int BR, int BC, int OV>
cudaError_t apply_q_t(const BandedArgs& a, cudaStream_t s) {
  if constexpr (OV == 0) {
    banded_apply_q_kernel<BR, BC, OV><<<groups_of(a), 32, 0, s>>>(a.packed, a.tau, a.y, a.x, a.nb, a.last_cols, a.group);
    return cudaGetLastError();
  }
  banded_chase_apply_kernel<BC, OV, true><<<1, 32, 0, s>>>(a.cvec, a.ctau, a.y, a.gy, a.nb, a.last_cols, a.group, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  banded_apply_q_kernel<BR, BC, OV><<<groups_of(a), 32, 0, s>>>(a.packed, a.tau, a.gy, a.x, a.nb, a.last_cols, a.group);
  return cudaGetLastError();
}
template <int BR, int BC, int OV>
cudaError_t backsolve_t(const BandedArgs& a, cudaStream_t s) {
  banded_backsolve_kernel<BC, OV><<<1, 32, 0, s>>>(a.rband, a.y, a.x, a.nb, a.last_cols);
  return cudaGetLastError();
}

#define X(BR, BC, OV) const BandedVTable kT_##BR##_##BC##_##OV = {BR, BC, OV, factor_t<BR, BC, OV>, apply_qt_t<BR, BC, OV>, apply_q_t<BR, BC, OV>, backsolve_t<BR, BC, OV>};
QRK_BANDED_SHAPES(X)
#undef X


// ---- the general window chain: the same four operations on BandedArgs::gen --------------------------------------------
template <typename K>
cudaError_t gen_opt_in(K kernel, size_t smem) {
  return smem > 48 * 1024 ? cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) : cudaSuccess;
}
cudaError_t gen_factor(const BandedArgs& a, cudaStream_t s) {
  GenArgs g = *a.gen;
  g.A_in = a.A_in; g.b = a.b; g.y = a.y; g.comp = a.comp;
  const size_t smem = gen_smem_factor(g.max_rows, g.max_cols);
  cudaError_t e = gen_opt_in(banded_generic_factor_kernel, smem);
  if (e != cudaSuccess) return e;
  banded_generic_factor_kernel<<<1, kGenThreads, smem, s>>>(g);
  return cudaGetLastError();
}
cudaError_t gen_apply_qt(const BandedArgs& a, cudaStream_t s) {
  const GenArgs g = *a.gen;
  const size_t smem = gen_smem_apply(g.max_rows, g.max_cols);
  cudaError_t e = gen_opt_in(banded_generic_apply_kernel<true>, smem);
  if (e != cudaSuccess) return e;
  banded_generic_apply_kernel<true><<<(unsigned)a.ncols, kGenThreads, smem, s>>>(g, a.b, a.ldb, a.y, a.ldy, a.comp, a.ldcomp, nullptr, 0);
  return cudaGetLastError();
}
cudaError_t gen_apply_q(const BandedArgs& a, cudaStream_t s) {
  const GenArgs g = *a.gen;
  const size_t smem = gen_smem_apply(g.max_rows, g.max_cols);
  cudaError_t e = gen_opt_in(banded_generic_apply_kernel<false>, smem);
  if (e != cudaSuccess) return e;
  banded_generic_apply_kernel<false><<<(unsigned)a.ncols, kGenThreads, smem, s>>>(g, nullptr, 0, a.y, a.ldy, a.comp, a.ldcomp, a.x, a.ldx);
  return cudaGetLastError();
}
cudaError_t gen_backsolve(const BandedArgs& a, cudaStream_t s) {
  banded_generic_backsolve_kernel<<<1, kGenThreads, 0, s>>>(*a.gen, a.y, a.x);
  return cudaGetLastError();
}
const BandedVTable kGenericVT = {0, 0, 0, gen_factor, gen_apply_qt, gen_apply_q, gen_backsolve};

}  // namespace

const BandedVTable* banded_generic_vtable() { return &kGenericVT; }
size_t banded_generic_smem_bytes(int max_rows, int max_cols) { return gen_smem_factor(max_rows, max_cols); }
cudaError_t banded_generic_export_r(const GenArgs& g, const int* d_col0, const int* d_outer, const int* d_inner, double* d_vals, cudaStream_t s) {
  banded_generic_export_r_kernel<<<148 * 4, 256, 0, s>>>(g, d_col0, d_outer, d_inner, d_vals);
  return cudaGetLastError();
}

const BandedVTable* banded_vtable(int br, int bc, int ov) {
#define X(BR, BC, OV) if (br == BR && bc == BC && ov == OV) return &kT_##BR##_##BC##_##OV;
  QRK_BANDED_SHAPES(X)
#undef X
  return nullptr;
}

int banded_launches_per_call() { return 2; }

long long banded_comp_rows(long long nb, int br, int bc, int ov, int group, int last_cols) {
  const long long groups = (nb + group - 1) / group;
  return nb * (long long)(ov + br - bc) + groups * ov + (bc - last_cols);
}

}  // namespace qrk
