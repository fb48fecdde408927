// Compile-and-link check of include/qrkit_b200/EigenAdapter.hpp against the mock Eigen / QRKit headers (tests/cpp/mock_*):
// every member of the three adapter classes is instantiated.  Without a CUDA device the program checks that compute() reports
// InvalidInput through info() (the adapter follows the reference's no-exception convention) and exits 77; with one it runs
// the three solvers on small problems and checks that x is recovered from a consistent system.
#include <cmath>
#include <cstdio>
#include "QRKitMock.h"
#include "qrkit_b200/EigenAdapter.hpp"

#ifndef QRKIT_B200_HAVE_EIGEN
#error "the mock <Eigen/Sparse> was not found by __has_include"
#endif

using namespace QRKit;
typedef Eigen::Matrix<double, 7, 2> Block7x2;
static_assert(sizeof(Block7x2) == 14 * sizeof(double), "fixed-size blocks are packed: std::vector<Block> is the block-COO value array");
typedef BlockDiagonalSparseQR_B200<Eigen::ColPivHouseholderQR<Block7x2> > DiagQR;
typedef BlockAngularSparseQR_B200<DiagQR, Eigen::ColPivHouseholderQR<Eigen::MatrixXd> > AngularQR;
typedef Eigen::SparseMatrix<double, Eigen::ColMajor, int> SpMat;
typedef BandedBlockedSparseQR_B200<SpMat, Eigen::HouseholderQR<Eigen::Matrix<double, 7, 4> >, 2> BandedQR;

template class QRKit::BlockDiagonalSparseQR_B200<Eigen::ColPivHouseholderQR<Block7x2> >;
template class QRKit::BlockAngularSparseQR_B200<DiagQR, Eigen::ColPivHouseholderQR<Eigen::MatrixXd> >;
template class QRKit::BandedBlockedSparseQR_B200<SpMat, Eigen::HouseholderQR<Eigen::Matrix<double, 7, 4> >, 2>;
static_assert(SparseQRUtils::HasRowsPermutation<DiagQR>::value && SparseQRUtils::HasRowsPermutation<AngularQR>::value &&
              SparseQRUtils::HasRowsPermutation<BandedQR>::value, "HasRowsPermutation trait (BlockDiagonalSparseQR.h:337-340)");

static double gen(unsigned a, unsigned b, unsigned c) { unsigned x = a * 2654435761u ^ b * 40503u ^ c * 69069u; x ^= x >> 13; x *= 1274126177u; x ^= x >> 16; return 0.5 + 4.5 * (x % 100000) / 100000.0; }

int main() {
  int ndev = 0;
  qrk_device_count(&ndev);
  const Eigen::Index nb = 64, r = 7, c = 2;
  SparseBlockDiagonal<Block7x2> A(nb * r, nb * c);
  for (Eigen::Index i = 0; i < nb; i++) {
    Block7x2 b;
    for (int j = 0; j < c; j++) for (int k = 0; k < r; k++) b.data()[j * r + k] = gen((unsigned)i, (unsigned)k, (unsigned)j);
    A.insertBack(b);
  }
  DiagQR diag;
  diag.compute(A);
  if (ndev < 1) {
    std::printf("no CUDA device: info() = %d\n", (int)diag.info());
    return diag.info() == Eigen::InvalidInput ? 77 : 1;
  }
  int failures = 0;
  Eigen::VectorXd x(nb * c), b(nb * r);
  for (Eigen::Index j = 0; j < nb * c; j++) x.data()[j] = gen(7, (unsigned)j, 1) - 2.5;
  for (Eigen::Index i = 0; i < nb; i++) for (int k = 0; k < r; k++) { double s = 0; for (int j = 0; j < c; j++) s += A[i].data()[j * r + k] * x.data()[i * c + j]; b.data()[i * r + k] = s; }
  Eigen::VectorXd xs = diag.solve(b);
  double err = 0, nrm = 0;
  for (Eigen::Index j = 0; j < nb * c; j++) { err += (xs.data()[j] - x.data()[j]) * (xs.data()[j] - x.data()[j]); nrm += x.data()[j] * x.data()[j]; }
  if (!(diag.info() == Eigen::Success && diag.rank() == nb * c && std::sqrt(err / nrm) <= 1e-10)) { std::printf("FAILED: BlockDiagonalSparseQR_B200 solve\n"); failures++; }
  if (diag.matrixR().nonZeros() != nb * c * (c + 1) / 2) { std::printf("FAILED: matrixR nnz\n"); failures++; }
  Eigen::MatrixXd Bm(nb * r, 1);
  for (Eigen::Index i = 0; i < nb * r; i++) Bm.data()[i] = b.data()[i];
  Eigen::MatrixXd back = diag.matrixQ() * (diag.matrixQ().transpose() * Bm);
  err = 0;
  for (Eigen::Index i = 0; i < nb * r; i++) err += (back.data()[i] - b.data()[i]) * (back.data()[i] - b.data()[i]);
  if (!(std::sqrt(err) <= 1e-10)) { std::printf("FAILED: Q Q^T b = b\n"); failures++; }
  std::printf(failures ? "FAILED (%d)\n" : "All passed.\n", failures);
  return failures ? 1 : 0;
}
