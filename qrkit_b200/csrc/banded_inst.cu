// banded_inst.cu — instantiations of the block-banded kernels.
#include "banded.cuh"
#include "banded_dispatch.hpp"

namespace qrk {
namespace {

// (block_rows, block_cols, overlap): BASELINE config 4 (16, 24, 16); the reference test pattern 7x4 with overlap 2
// (test/test-qrkit.cpp:63-96) and the non-overlapping 7x2 pattern (:101-131); a few more for coverage.
#define QRK_BANDED_SHAPES(X) X(16, 24, 16) X(7, 4, 2) X(7, 2, 0) X(8, 8, 4) X(12, 8, 4) X(4, 6, 4)

template <int BR, int BC, int OV>
cudaError_t factor_t(const BandedArgs& a, cudaStream_t s) {
  banded_factor_kernel<BR, BC, OV><<<1, 32, 0, s>>>(a.A_in, a.packed, a.tau, a.rband, a.b, a.y, a.ycomp, a.nb, a.last_cols);
  return cudaGetLastError();
}
template <int BR, int BC, int OV>
cudaError_t apply_qt_t(const BandedArgs& a, cudaStream_t s) {
  banded_apply_qt_kernel<BR, BC, OV><<<1, 32, 0, s>>>(a.packed, a.tau, a.b, a.y, a.ycomp, a.nb, a.last_cols);
  return cudaGetLastError();
}
template <int BR, int BC, int OV>
cudaError_t backsolve_t(const BandedArgs& a, cudaStream_t s) {
  banded_backsolve_kernel<BC, OV><<<1, 32, 0, s>>>(a.rband, a.y, a.x, a.nb, a.last_cols);
  return cudaGetLastError();
}

#define X(BR, BC, OV) const BandedVTable kT_##BR##_##BC##_##OV = {BR, BC, OV, factor_t<BR, BC, OV>, apply_qt_t<BR, BC, OV>, backsolve_t<BR, BC, OV>};
QRK_BANDED_SHAPES(X)
#undef X

}  // namespace

const BandedVTable* banded_vtable(int br, int bc, int ov) {
#define X(BR, BC, OV) if (br == BR && bc == BC && ov == OV) return &kT_##BR##_##BC##_##OV;
  QRK_BANDED_SHAPES(X)
#undef X
  return nullptr;
}

}  // namespace qrk
