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
  try {
    solver.compute(A);
  } catch (const std::exception& e) {
    if (ndev >= 1) throw;
    // no device: compute() must refuse loudly (there is no CPU fallback), never return as if it had factorised
    std::printf("no CUDA device: compute() threw \"%s\"; info() = %d\n", e.what(), (int)solver.info());
    return (std::string(e.what()).find("no CPU fallback") != std::string::npos && solver.info() == InvalidInput) ? 77 : 1;
  }
  if (ndev < 1) {
    std::printf("FAILED: compute() returned without a CUDA device\n");
    return 1;
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

  // block angular with a wide border (m2 = 24) and the unpivoted right solver (test/test-qrkit.cpp:53-56, 294-327)
  {
    const Index nw = 200, m2 = 24;
    SparseBlockDiagonal<Block7x2> L1(nw * 7, nw * 2);
    MatrixXd B2(nw * 7, m2);
    for (Index i = 0; i < nw; i++) {
      Block7x2 blk;
      for (int j = 0; j < 2; j++) for (int k = 0; k < 7; k++) blk(k, j) = synth(21, i, k, j);
      L1.insertBack(blk);
    }
    for (Index j = 0; j < m2; j++) for (Index i = 0; i < nw * 7; i++) B2(i, j) = synth(23, 1, i, j);
    BlockMatrix1x2<SparseBlockDiagonal<Block7x2>, MatrixXd> Jw(L1, B2);
    VectorXd xw(nw * 2 + m2), bw(nw * 7, 0.0);
    for (size_t j = 0; j < xw.size(); j++) xw[j] = synth(27, j, 0, 0, -1.0, 1.0);
    for (Index i = 0; i < nw; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 7; k++) bw[i * 7 + k] += L1[i](k, j) * xw[i * 2 + j];
    for (Index j = 0; j < m2; j++) for (Index i = 0; i < nw * 7; i++) bw[i] += B2(i, j) * xw[nw * 2 + j];
    BlockAngularSparseQR<ColPivHouseholderQR<Block7x2>> wide(Jw);
    CHECK(wide.info() == Success && wide.rank() == nw * 2 + m2, "wide border (24 columns): info / rank");
    CHECK(rel(wide.solve(bw), xw) <= 1e-10, "wide border: solve(b) recovers x  (1e-10)");
    BlockAngularSparseQR<ColPivHouseholderQR<Block7x2>, BlockedThinDenseQR<MatrixXd, 2>> thin(Jw);
    CHECK(thin.rank() == nw * 2 + m2, "BlockedThinDenseQR right solver: rank = cols");
    bool ident = true;
    const auto& Pw = thin.colsPermutation();
    for (Index j = 0; j < m2; j++) ident = ident && Pw.indices()[nw * 2 + j] == nw * 2 + j;
    CHECK(ident, "BlockedThinDenseQR right solver: P2 = identity");
    CHECK(rel(thin.solve(bw), xw) <= 1e-10, "BlockedThinDenseQR right solver: solve(b) recovers x  (1e-10)");
  }

  // block banded: the overlapping 7x4 / overlap 2 pattern of test/test-qrkit.cpp:63-96 (dense slabs), x recovered (:255)
  {
    const Index nbb = 128;
    std::vector<double> slabs((size_t)nbb * 7 * 4);
    for (Index k = 0; k < nbb; k++) for (int j = 0; j < 4; j++) for (int i = 0; i < 7; i++) slabs[(size_t)(k * 4 + j) * 7 + i] = synth(31, k, i, j);
    BandedBlockedSparseQR<7, 4, 2> band;
    band.compute(slabs, nbb);
    CHECK(band.info() == Success && band.rows() == nbb * 7 && band.cols() == (nbb - 1) * 2 + 4 && band.rank() == band.cols(), "banded: info / rows / cols / rank");
    VectorXd xb((size_t)band.cols()), bb((size_t)band.rows(), 0.0);
    for (size_t j = 0; j < xb.size(); j++) xb[j] = synth(33, j, 0, 0, -1.0, 1.0);
    for (Index k = 0; k < nbb; k++) for (int j = 0; j < 4; j++) for (int i = 0; i < 7; i++) bb[k * 7 + i] += slabs[(size_t)(k * 4 + j) * 7 + i] * xb[k * 2 + j];
    CHECK(rel(band.solve(bb), xb) <= 1e-10, "banded: solve(b) recovers x  (1e-10)");
    BandedBlockedSparseQR<7, 4, 2> band2;
    CHECK(rel(band2.computeAndSolve(slabs, nbb, bb), xb) <= 1e-10, "banded: computeAndSolve recovers x  (1e-10)");
    const auto& Rb = band.matrixR();
    bool upper = true;
    for (Index j = 0; j < band.cols(); j++) for (int p = Rb.outer[j]; p < Rb.outer[j + 1]; p++) upper = upper && (Rb.inner[p] <= j || Rb.values[p] == 0.0);
    CHECK(upper, "banded: matrixR() is upper triangular");
    // matrixQ() * (matrixQ().transpose() * b) on the thin part = A x for a consistent b (Q1 Q1^T projects on range(A))
    VectorXd qqt = band.applyQ(band.applyQt(bb));
    CHECK(rel(qqt, bb) <= 1e-11, "banded: matrixQ() * (matrixQ().transpose() * b) = b for b in range(A)  (1e-11)");
    // the same banded block as the LEFT block of a block-angular matrix (the solver pair of test/test-qrkit.cpp:44-48),
    // dense border of 20 columns: x recovered from a consistent system (:289)
    const int mb = 20;
    MatrixXd Bb(band.rows(), mb);
    for (Index i = 0; i < band.rows(); i++) for (int j = 0; j < mb; j++) Bb(i, j) = synth(37, i, j, 0);
    VectorXd xa((size_t)(band.cols() + mb)), ba((size_t)band.rows(), 0.0);
    for (size_t j = 0; j < xa.size(); j++) xa[j] = synth(39, j, 0, 0, -1.0, 1.0);
    for (Index k = 0; k < nbb; k++) for (int j = 0; j < 4; j++) for (int i = 0; i < 7; i++) ba[k * 7 + i] += slabs[(size_t)(k * 4 + j) * 7 + i] * xa[k * 2 + j];
    for (Index i = 0; i < band.rows(); i++) for (int j = 0; j < mb; j++) ba[i] += Bb(i, j) * xa[band.cols() + j];
    BlockAngularBandedSparseQR<7, 4, 2> ab;
    ab.compute(slabs, nbb, Bb);
    CHECK(ab.info() == Success && ab.rows() == band.rows() && ab.cols() == band.cols() + mb && ab.rank() == ab.cols(), "banded-left angular: info / rows / cols / rank");
    CHECK(rel(ab.solve(ba), xa) <= 1e-10, "banded-left angular: solve(b) recovers x  (1e-10)");
    BlockAngularBandedSparseQR<7, 4, 2, BlockedThinDenseQR<MatrixXd, 2>> ab2;
    CHECK(rel(ab2.computeAndSolve(slabs, nbb, Bb, ba), xa) <= 1e-10, "banded-left angular, unpivoted right solver: computeAndSolve recovers x  (1e-10)");
  }
  // BandedBlockedSparseQR::compute(const SparseMatrix&) on a general banded matrix with SHUFFLED rows (test/test-qrkit.cpp:63-96,
  // 208-258): row ordering + block detection + the general window chain; b is permuted with rowsPermutation() before solve (:235)
  {
    const Index np = 40, nbk = np / 2, nr = 7 * nbk;
    std::vector<std::vector<std::pair<Index, double>>> colsv((size_t)np);       // column -> (row, value), rows shuffled
    std::vector<Index> shuffle((size_t)nr);
    for (Index i = 0; i < nr; i++) shuffle[(size_t)i] = (i * 37 + 11) % nr;     // a fixed permutation (37 and nr = 140 are coprime)
    for (Index i = 0; i < nbk; i++)
      for (Index j = 2 * i; j < 2 * i + 2; j++) {
        for (int k = 0; k < 7; k++) colsv[(size_t)j].push_back({shuffle[(size_t)(i * 7 + k)], synth(61, i, k, j)});
        if (j < np - 2) colsv[(size_t)j + 2].push_back({shuffle[(size_t)(i * 7 + 6)], synth(63, i, 6, j)});
      }
    SparseMatrix<ColMajor> Sg;
    Sg.m_rows = nr; Sg.m_cols = np; Sg.outer.assign((size_t)np + 1, 0);
    for (Index j = 0; j < np; j++) {
      std::sort(colsv[(size_t)j].begin(), colsv[(size_t)j].end());
      for (auto& e : colsv[(size_t)j]) { Sg.inner.push_back((StorageIndex)e.first); Sg.values.push_back(e.second); }
      Sg.outer[(size_t)j + 1] = (StorageIndex)Sg.inner.size();
    }
    BandedBlockedSparseQR<7, 4, 2> gen;
    gen.compute(Sg);
    CHECK(gen.info() == Success && gen.rows() == nr && gen.cols() == np && gen.rank() == np, "banded, general sparse input: info / rows / cols / rank");
    VectorXd xg((size_t)np), bg((size_t)nr, 0.0);
    for (size_t j = 0; j < xg.size(); j++) xg[j] = synth(65, j, 0, 0, -1.0, 1.0);
    for (Index j = 0; j < np; j++) for (StorageIndex p = Sg.outer[(size_t)j]; p < Sg.outer[(size_t)j + 1]; p++) bg[(size_t)Sg.inner[(size_t)p]] += Sg.values[(size_t)p] * xg[(size_t)j];
    const VectorXd pb = gen.rowsPermutation() * bg;
    CHECK(rel(gen.solve(pb), xg) <= 1e-10, "banded, general sparse input with shuffled rows: solve(P b) recovers x  (1e-10)");
    const VectorXd qt = gen.applyQtFull(pb);
    double n1 = 0, n2 = 0;
    for (size_t i = 0; i < qt.size(); i++) { n1 += qt[i] * qt[i]; n2 += pb[i] * pb[i]; }
    CHECK(std::fabs(std::sqrt(n1) - std::sqrt(n2)) <= 1e-13 * std::sqrt(n2) && rel(gen.applyQFull(qt), pb) <= 1e-13, "banded, general path: n x n Q is orthogonal (||Q^T v|| = ||v||, Q Q^T v = v)");
  }
  // rank-revealing right solver (BlockedThinSparseQR, test/test-qrkit.cpp:54-57)
  {
    const Index nbt = 64;
    SparseBlockDiagonal<Block7x2> At(nbt * r, nbt * c);
    for (Index i = 0; i < nbt; i++) { Block7x2 bb; for (int j = 0; j < c; j++) for (int k = 0; k < r; k++) bb(k, j) = synth(71, i, k, j); At.insertBack(bb); }
    MatrixXd Bt(nbt * r, 10);
    for (Index i = 0; i < nbt * r; i++) for (int j = 0; j < 10; j++) Bt(i, j) = synth(73, i, j, 0);
    VectorXd xt((size_t)(nbt * c + 10)), bt((size_t)(nbt * r), 0.0);
    for (size_t j = 0; j < xt.size(); j++) xt[j] = synth(75, j, 0, 0, -1.0, 1.0);
    for (Index i = 0; i < nbt; i++) for (int j = 0; j < c; j++) for (int k = 0; k < r; k++) bt[(size_t)(i * r + k)] += At[i](k, j) * xt[(size_t)(i * c + j)];
    for (Index i = 0; i < nbt * r; i++) for (int j = 0; j < 10; j++) bt[(size_t)i] += Bt(i, j) * xt[(size_t)(nbt * c + j)];
    BlockMatrix1x2<SparseBlockDiagonal<Block7x2>, MatrixXd> Mt(At, Bt);
    BlockAngularSparseQR<ColPivHouseholderQR<Block7x2>, BlockedThinSparseQR<SparseMatrix<ColMajor>, 2>> ts(Mt);
    CHECK(ts.info() == Success && ts.rank() == nbt * c + 10, "BlockedThinSparseQR right solver: info / rank");
    CHECK(rel(ts.solve(bt), xt) <= 1e-10, "BlockedThinSparseQR right solver: solve(b) recovers x  (1e-10)");
  }
  // one host process, several GPUs (two shards on device 0 when the box has one GPU): contiguous block ranges per shard
  {
    const std::vector<int> devs = {0, ndev > 1 ? 1 : 0, 0};
    const Index nbs = 1001;                                     // odd: ragged last shard
    SparseBlockDiagonal<Block7x2> As(nbs * r, nbs * c);
    for (Index i = 0; i < nbs; i++) { Block7x2 bb; for (int j = 0; j < c; j++) for (int k = 0; k < r; k++) bb(k, j) = synth(41, i, k, j); As.insertBack(bb); }
    VectorXd bs((size_t)(nbs * r));
    for (size_t i = 0; i < bs.size(); i++) bs[i] = synth(43, i, 0, 0, -1.0, 1.0);
    DiagQR one;
    const VectorXd x_one = one.computeAndSolve(As, bs);
    ShardedBlockDiagonalSparseQR<ColPivHouseholderQR<Block7x2>> sh(devs);
    const VectorXd x_sh = sh.computeAndSolve(As, bs);
    CHECK(x_sh == x_one, "sharded block-diagonal computeAndSolve over 3 shards: bit-identical to one handle");
    sh.compute(As);
    CHECK(rel(sh.solve(bs), x_one) <= 1e-13 && sh.rank() == nbs * c, "sharded block-diagonal compute() + solve() (1e-13), rank");
    const auto Ps = sh.colsPermutation(); const auto& Po = one.colsPermutation();
    bool same = true;
    for (Index j = 0; j < nbs * c; j++) same = same && Ps.indices()[(size_t)j] == Po.indices()[(size_t)j];
    CHECK(same, "sharded colsPermutation() = the single-handle permutation (global column indices)");
    // block angular, 5 border columns: the triangles cross over peer memory inside the TSQR root kernel
    typedef Matrix<2, 1> Block2x1;
    const Index np = 4000;
    SparseBlockDiagonal<Block2x1> Ja(np * 2, np);
    for (Index i = 0; i < np; i++) { Block2x1 bb; bb(0, 0) = synth(51, i, 0, 0); bb(1, 0) = synth(51, i, 1, 0); Ja.insertBack(bb); }
    MatrixXd Jb(np * 2, 5);
    for (Index i = 0; i < np * 2; i++) for (int j = 0; j < 5; j++) Jb(i, j) = synth(53, i, j, 0);
    VectorXd ba((size_t)(np * 2));
    for (size_t i = 0; i < ba.size(); i++) ba[i] = synth(55, i, 0, 0, -1.0, 1.0);
    BlockMatrix1x2<SparseBlockDiagonal<Block2x1>, MatrixXd> Ma(Ja, Jb);
    BlockAngularSparseQR<ColPivHouseholderQR<Block2x1>> a1;
    const VectorXd xa1 = a1.computeAndSolve(Ma, ba);
    ShardedBlockAngularSparseQR<ColPivHouseholderQR<Block2x1>> ash(std::vector<int>{0, ndev > 1 ? 1 : 0});
    const VectorXd xas = ash.computeAndSolve(Ma, ba);
    CHECK(rel(xas, xa1) <= 1e-10 && ash.sharedParametersIdentical(), "sharded block-angular computeAndSolve (fused peer exchange): x (1e-10), x2 bit-identical on the shards");
    CHECK(ash.rank() == np + 5, "sharded block-angular rank");
  }
  std::printf(failures ? "FAILED (%d)\n" : "All passed.\n", failures);
  return failures ? 1 : 0;
}
