// test_facade.cpp — the reference's block-diagonal test (test/test-qrkit.cpp:167-206: 256 blocks of 7x2,
// ColPivHouseholderQR per block) and a small block-angular case (shape of :260-292 with a block-diagonal left
// solver), written against include/qrkit_b200/QRKit.hpp exactly as the reference's test is written against
// QRKit: Q*R = A*P, Q^T*(A*P) = R, x recovered.  Tolerances are the north-star ones, not the reference's 1e-6.
// Exit code: 0 = all passed, 1 = a property failed, 77 = no CUDA device (nothing computed: there is no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdint>

#include "qrkit_b200/QRKit.hpp"

using namespace QRKit_b200;

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
static double synth(uint64_t seed, uint64_t block, uint64_t row, uint64_t col, double lo = 0.5, double hi = 5.0) {
  const uint64_t u = splitmix64(seed ^ (block << 20) ^ (row << 10) ^ col);
  return lo + (hi - lo) * (double)(u >> 11) * (1.0 / 9007199254740992.0);
}
static double rel(const VectorXd& a, const VectorXd& b) {
  double n = 0, d = 0;
  for (size_t i = 0; i < a.size(); i++) { n += (a[i] - b[i]) * (a[i] - b[i]); d += b[i] * b[i]; }
  return std::sqrt(n) / std::sqrt(d > 0 ? d : 1);
}
static int failures = 0;
#define CHECK(cond, what) do { if (!(cond)) { std::printf("FAILED: %s\n", what); failures++; } else std::printf("Passed: %s\n", what); } while (0)

typedef Matrix<7, 2> Block7x2;
typedef BlockDiagonalSparseQR<ColPivHouseholderQR<Block7x2>> DiagQR;

int main() {
  int ndev = 0;
  qrk_device_count(&ndev);
  const Index nb = 256, r = 7, c = 2, rows = nb * r, cols = nb * c;
  SparseBlockDiagonal<Block7x2> A(rows, cols);
  for (Index i = 0; i < nb; i++) {
    Block7x2 b;
    for (int j = 0; j < c; j++) for (int k = 0; k < r; k++) b(k, j) = synth(0x51524B49, i, k, j);
    A.insertBack(b);
  }
  DiagQR solver;
  solver.compute(A);
  if (ndev < 1) {
    std::printf("no CUDA device: info() = %d (%s)\n", (int)solver.info(), solver.lastErrorMessage().c_str());
    return solver.info() == InvalidInput ? 77 : 1;
  }
  CHECK(solver.info() == Success, "info() == Success");
  CHECK(solver.rows() == rows && solver.cols() == cols && solver.rank() == cols, "rows / cols / rank");
  const auto& R = solver.matrixR();
  const auto Q = solver.matrixQ().toSparse();
  const auto& P = solver.colsPermutation();
  // Q * R = A * P  and  Q^T * (A * P) = R, evaluated block column by block column
  double err_qr = 0, err_qta = 0, nrm = 0;
  for (Index j = 0; j < cols; j++) {
    VectorXd ap(rows, 0.0), rcol(rows, 0.0);
    const Index src = P.indices()[j], blk = src / c;
    for (int k = 0; k < r; k++) ap[blk * r + k] = A[blk](k, src % c);
    for (int p = R.outer[j]; p < R.outer[j + 1]; p++) rcol[R.inner[p]] = R.values[p];
    const VectorXd qr = solver.matrixQ() * rcol, qta = solver.matrixQ().transpose() * ap;
    for (Index i = 0; i < rows; i++) {
      err_qr += (qr[i] - ap[i]) * (qr[i] - ap[i]); err_qta += (qta[i] - rcol[i]) * (qta[i] - rcol[i]); nrm += ap[i] * ap[i];
    }
  }
  CHECK(std::sqrt(err_qr / nrm) <= 1e-13, "Q * R = A * P  (1e-13)");
  CHECK(std::sqrt(err_qta / nrm) <= 1e-13, "Q^T * (A * P) = R  (1e-13)");
  CHECK(Q.nonZeros() == nb * r * r && R.nonZeros() == nb * c * (c + 1) / 2, "nnz of explicit Q and R");
  // solve: b = A * x_true
  VectorXd x_true(cols), b(rows, 0.0);
  for (Index j = 0; j < cols; j++) x_true[j] = synth(7, j, 0, 0, -1.0, 1.0);
  for (Index i = 0; i < nb; i++) for (int j = 0; j < c; j++) for (int k = 0; k < r; k++) b[i * r + k] += A[i](k, j) * x_true[i * c + j];
  CHECK(rel(solver.solve(b), x_true) <= 1e-10, "solve(b) recovers x  (1e-10)");
  DiagQR fused;
  CHECK(rel(fused.computeAndSolve(A, b), x_true) <= 1e-10, "computeAndSolve(A, b) recovers x  (1e-10)");

  // block angular: 2x1 blocks + 5 dense border columns (the ellipse-fit shape, bench/bench_sparse_qr_extra.cpp:100-114)
  typedef Matrix<2, 1> Block2x1;
  const Index n = 1000;
  SparseBlockDiagonal<Block2x1> J1(2 * n, n);
  MatrixXd J2(2 * n, 5);
  for (Index i = 0; i < n; i++) {
    Block2x1 blk; blk(0, 0) = synth(11, i, 0, 0); blk(1, 0) = synth(11, i, 1, 0);
    J1.insertBack(blk);
  }
  for (Index j = 0; j < 5; j++) for (Index i = 0; i < 2 * n; i++) J2(i, j) = synth(13, 1, i, j, -1.0, 1.0);
  BlockMatrix1x2<SparseBlockDiagonal<Block2x1>, MatrixXd> J(J1, J2);
  BlockAngularSparseQR<ColPivHouseholderQR<Block2x1>> ang(J);
  CHECK(ang.info() == Success && ang.rank() == n + 5 && ang.cols() == n + 5, "block angular: info / rank / cols");
  VectorXd xa(n + 5), ba(2 * n, 0.0);
  for (Index j = 0; j < n + 5; j++) xa[j] = synth(17, j, 0, 0, -1.0, 1.0);
  for (Index i = 0; i < n; i++) for (int k = 0; k < 2; k++) ba[2 * i + k] += J1[i](k, 0) * xa[i];
  for (Index j = 0; j < 5; j++) for (Index i = 0; i < 2 * n; i++) ba[i] += J2(i, j) * xa[n + j];
  CHECK(rel(ang.solve(ba), xa) <= 1e-10, "block angular: solve(b) recovers x  (1e-10)");
  const auto& Ra = ang.matrixR();
  CHECK(Ra.nonZeros() == n + 5 * n + 15, "block angular: nnz(R) = nnz(R1) + m1*m2 + m2(m2+1)/2");
  std::printf(failures ? "FAILED (%d)\n" : "All passed.\n", failures);
  return failures ? 1 : 0;
}
