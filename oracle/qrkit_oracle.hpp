// qrkit_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A dependency-free C++17 restatement of the structured-sparse QR hot path of
// jasvob/QRKit, used only as the checker for the CUDA path:
//   * tests/            compare the CUDA results with these functions,
//   * __graft_entry__.smoke()   does the same on one small case,
//   * bench.py          times it as the `cpu_baseline` / `--impl reference` leg.
// Nothing under qrkit_b200/ (the product) includes, links or calls this file.
//
// PARITY UNPINNED.  The reference (header-only C++ on top of Eigen) cannot be
// compiled in this image because Eigen — an un-vendored dependency required as
// `find_package(Eigen3 3.3 REQUIRED NO_MODULE)` (reference CMakeLists.txt:5), no
// version pinned beyond ">= 3.3" — is absent and there is no network.  The
// reference's own tests hold no golden vectors (only properties at 1e-6,
// test/test.h:31).  So this oracle is pinned by
//   (1) the reference's properties Q·R = A·P, Qᵀ·A·P = R, x recovered
//       (test/test-qrkit.cpp:201-203, 251-255, 289), asserted at 1e-13/1e-10,
//   (2) an independent LAPACK cross-check (scipy dgeqrf/dgeqp3) of R up to row
//       sign and of the pivot order, and numpy lstsq for x
// (tests/test_oracle.py).  It is NOT pinned against an Eigen binary.
//
// What is restated, and from where:
//   QRKit side (citable, file:line relative to /root/reference):
//     BlockDiagonalSparseQR::factorize      src/QRKit/BlockDiagonalSparseQR.h:415-547
//     BlockDiagonalSparseQR::_solve_impl    src/QRKit/BlockDiagonalSparseQR.h:258-280
//     BlockAngularSparseQR::factorize       src/QRKit/BlockAngularSparseQR.h:459-514
//       solveRightBlock (dense)             src/QRKit/BlockAngularSparseQR.h:361-369
//       makeR (dense)                       src/QRKit/BlockAngularSparseQR.h:285-308
//       Qᵀ·v                                src/QRKit/BlockAngularSparseQR.h:607-625
//     BandedBlockedSparseQR::factorize      src/QRKit/BandedBlockedSparseQR.h:443-519
//     SparseBlockYTY sequence product       src/QRKit/SparseBlockYTY.h:102-139
//     fromBlockDiagonalPattern / fromBlockBandedPattern / mergeBlocks
//                                           src/QRKit/SparseQRUtils.h:255-385
//     BlockedThinDenseQR                    src/QRKit/BlockedThinDenseQR.h:104-176,
//                                           src/QRKit/BlockedThinQRBase.h:308-333
//   Eigen side (third-party, NOT in the reference tree; restated from Eigen 3.3's
//   published algorithm): makeHouseholder, HouseholderQR (unblocked + 48-column
//   panels), ColPivHouseholderQR (LAWN-176 norm downdating, first-max pivot),
//   HouseholderSequence -> dense Q, make_block_householder_triangular_factor,
//   SparseMatrix::setFromTriplets (sorted, duplicates summed), sparse upper
//   triangular solve, PermutationMatrix products.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <numeric>
#include <vector>

namespace qrk_oracle {

// ----------------------------------------------------------------------------------------
// Dense column-major matrix (stands in for Eigen::MatrixXd)
// ----------------------------------------------------------------------------------------
struct Dense {
  int rows = 0, cols = 0;
  std::vector<double> v;
  Dense() {}
  Dense(int r, int c) : rows(r), cols(c), v((size_t)r * c, 0.0) {}
  double& operator()(int i, int j) { return v[(size_t)j * rows + i]; }
  double operator()(int i, int j) const { return v[(size_t)j * rows + i]; }
  double* col(int j) { return v.data() + (size_t)j * rows; }
  const double* col(int j) const { return v.data() + (size_t)j * rows; }
  static Dense identity(int r, int c) {
    Dense d(r, c);
    for (int i = 0; i < std::min(r, c); i++) d(i, i) = 1.0;
    return d;
  }
  Dense block(int r0, int c0, int nr, int nc) const {
    Dense d(nr, nc);
    for (int j = 0; j < nc; j++)
      for (int i = 0; i < nr; i++) d(i, j) = (*this)(r0 + i, c0 + j);
    return d;
  }
};

// ----------------------------------------------------------------------------------------
// Eigen::MatrixBase::makeHouseholderInPlace   [Eigen 3.3 Householder/Householder.h]
//   x: n-vector, contiguous.  On exit x[1:] holds the essential part; returns tau, beta.
//   H = I - tau * v v^T, v = [1; ess], H x = beta e1, beta = -sign(x0) * ||x|| (x0 >= 0 -> beta < 0).
// ----------------------------------------------------------------------------------------
inline void make_householder_inplace(double* x, int n, double& tau, double& beta) {
  double tailSq = 0.0;
  for (int i = 1; i < n; i++) tailSq += x[i] * x[i];
  const double c0 = x[0];
  const double tol = DBL_MIN;
  if (n == 1 || tailSq <= tol) {
    tau = 0.0;
    beta = c0;
    for (int i = 1; i < n; i++) x[i] = 0.0;
  } else {
    beta = std::sqrt(c0 * c0 + tailSq);
    if (c0 >= 0.0) beta = -beta;
    const double d = c0 - beta;
    for (int i = 1; i < n; i++) x[i] /= d;
    tau = (beta - c0) / beta;
  }
}

// Eigen::MatrixBase::applyHouseholderOnTheLeft on a (n x ncols) col-major panel with leading
// dimension ld: panel = (I - tau v v^T) panel, v = [1; ess] (ess has n-1 entries).
inline void apply_householder_left(double* panel, int ld, int n, int ncols, const double* ess, double tau) {
  if (n == 1) {
    for (int j = 0; j < ncols; j++) panel[(size_t)j * ld] *= (1.0 - tau);
    return;
  }
  if (tau == 0.0) return;
  for (int j = 0; j < ncols; j++) {
    double* c = panel + (size_t)j * ld;
    double tmp = 0.0;
    for (int i = 1; i < n; i++) tmp += ess[i - 1] * c[i];
    tmp += c[0];
    c[0] -= tau * tmp;
    for (int i = 1; i < n; i++) c[i] -= tau * ess[i - 1] * tmp;
  }
}

// Eigen::internal::householder_qr_inplace_unblocked on an (r x c) panel, leading dimension ld.
inline void householder_qr_unblocked(double* A, int ld, int r, int c, double* tau) {
  const int size = std::min(r, c);
  for (int k = 0; k < size; k++) {
    const int remRows = r - k, remCols = c - k - 1;
    double beta;
    double* x = A + (size_t)k * ld + k;
    make_householder_inplace(x, remRows, tau[k], beta);
    x[0] = beta;
    if (remCols > 0) apply_householder_left(A + (size_t)(k + 1) * ld + k, ld, remRows, remCols, x + 1, tau[k]);
  }
}

// Eigen::internal::make_block_householder_triangular_factor: T (n x n upper) such that
// H_0 ... H_{n-1} = I - V T V^T, V unit lower trapezoidal (rows x n) — only the strictly lower part
// of V is read, the diagonal is taken as 1.
inline Dense block_householder_t_factor(const double* V, int ldv, int rows, int n, const double* tau) {
  Dense T(n, n);
  for (int i = n - 1; i >= 0; --i) {
    const int rt = n - i - 1;
    if (rt > 0) {
      // T(i, i+1+jj) = -tau_i * V[i+1:, i]^T * unitLower(V[i+1:, i+1:])[:, jj]
      for (int jj = 0; jj < rt; jj++) {
        const int col = i + 1 + jj;
        // unit lower column `col` restricted to rows i+1..rows-1: zero above row `col`, one at `col`.
        double s = V[(size_t)i * ldv + col];  // * 1 (unit diagonal)
        for (int rr = col + 1; rr < rows; rr++) s += V[(size_t)i * ldv + rr] * V[(size_t)col * ldv + rr];
        T(i, col) = -tau[i] * s;
      }
      for (int j = n - 1; j > i; --j) {
        const double z = T(i, j);
        T(i, j) = z * T(j, j);
        for (int l = j + 1; l < n; l++) T(i, l) += z * T(j, l);
      }
    }
    T(i, i) = tau[i];
  }
  return T;
}

// Eigen::internal::apply_block_householder_on_the_left(mat, vectors, hCoeffs, forward=false):
//   mat = (I - V T V^T)^T mat = mat - V T^T V^T mat.
inline void apply_block_householder_left_adjoint(double* M, int ldm, int rows, int mcols, const double* V, int ldv,
                                                 int nb, const double* tau) {
  Dense T = block_householder_t_factor(V, ldv, rows, nb, tau);
  std::vector<double> tmp((size_t)nb * mcols), tmp2((size_t)nb * mcols);
  auto Vat = [&](int i, int j) -> double { return i == j ? 1.0 : (i < j ? 0.0 : V[(size_t)j * ldv + i]); };
  for (int j = 0; j < mcols; j++)
    for (int k = 0; k < nb; k++) {
      double s = 0.0;
      for (int i = k; i < rows; i++) s += Vat(i, k) * M[(size_t)j * ldm + i];
      tmp[(size_t)j * nb + k] = s;
    }
  for (int j = 0; j < mcols; j++)
    for (int k = 0; k < nb; k++) {  // T^T (lower) * tmp
      double s = 0.0;
      for (int l = 0; l <= k; l++) s += T(l, k) * tmp[(size_t)j * nb + l];
      tmp2[(size_t)j * nb + k] = s;
    }
  for (int j = 0; j < mcols; j++)
    for (int i = 0; i < rows; i++) {
      double s = 0.0;
      for (int k = 0; k < nb && k <= i; k++) s += Vat(i, k) * tmp2[(size_t)j * nb + k];
      M[(size_t)j * ldm + i] -= s;
    }
}

// Eigen::HouseholderQR<MatrixXd>::compute: householder_qr_inplace_blocked with maxBlockSize = 48.
// For min(r,c) <= 48 this is exactly the unblocked algorithm.
inline void householder_qr(double* A, int r, int c, double* tau, int maxBlockSize = 48) {
  const int size = std::min(r, c);
  const int blockSize = std::min(maxBlockSize, size);
  if (blockSize <= 0) return;
  for (int k = 0; k < size; k += blockSize) {
    const int bs = std::min(size - k, blockSize);
    const int tcols = c - k - bs;
    const int brows = r - k;
    double* A11 = A + (size_t)k * r + k;
    householder_qr_unblocked(A11, r, brows, bs, tau + k);
    if (tcols > 0)
      apply_block_householder_left_adjoint(A + (size_t)(k + bs) * r + k, r, brows, tcols, A11, r, bs, tau + k);
  }
}

// Eigen::ColPivHouseholderQR<>::computeInPlace (Eigen >= 3.3: LAWN-176 norm downdating, first-max pivot).
// perm[j] = index of the original column that ends up at position j (A*P = Q*R, P.indices()(j) = perm[j]).
// Returns nonzero_pivots.
inline int colpiv_householder_qr(double* A, int r, int c, double* tau, int* perm) {
  const int size = std::min(r, c);
  std::vector<double> upd(c), dir(c);
  std::vector<int> transp(size);
  auto colnorm = [&](int j, int from) {
    double s = 0.0;
    for (int i = from; i < r; i++) s += A[(size_t)j * r + i] * A[(size_t)j * r + i];
    return std::sqrt(s);
  };
  for (int j = 0; j < c; j++) upd[j] = dir[j] = colnorm(j, 0);
  double maxn = 0.0;
  for (int j = 0; j < c; j++) maxn = std::max(maxn, upd[j]);
  const double eps = DBL_EPSILON;
  const double threshold_helper = (maxn * eps) * (maxn * eps) / double(r);
  const double norm_downdate_threshold = std::sqrt(eps);
  int nonzero_pivots = size;
  for (int k = 0; k < size; k++) {
    int big = k;
    double bigv = upd[k];
    for (int j = k + 1; j < c; j++)
      if (upd[j] > bigv) { bigv = upd[j]; big = j; }  // strict '>' keeps the FIRST maximum
    const double big_sq = bigv * bigv;
    if (nonzero_pivots == size && big_sq < threshold_helper * double(r - k)) nonzero_pivots = k;
    transp[k] = big;
    if (k != big) {
      for (int i = 0; i < r; i++) std::swap(A[(size_t)k * r + i], A[(size_t)big * r + i]);
      std::swap(upd[k], upd[big]);
      std::swap(dir[k], dir[big]);
    }
    double beta;
    double* x = A + (size_t)k * r + k;
    make_householder_inplace(x, r - k, tau[k], beta);
    x[0] = beta;
    if (c - k - 1 > 0) apply_householder_left(A + (size_t)(k + 1) * r + k, r, r - k, c - k - 1, x + 1, tau[k]);
    for (int j = k + 1; j < c; j++) {
      if (upd[j] != 0.0) {
        double temp = std::fabs(A[(size_t)j * r + k]) / upd[j];
        temp = (1.0 + temp) * (1.0 - temp);
        temp = temp < 0.0 ? 0.0 : temp;
        const double q = upd[j] / dir[j];
        const double temp2 = temp * (q * q);
        if (temp2 <= norm_downdate_threshold) {
          dir[j] = colnorm(j, k + 1);
          upd[j] = dir[j];
        } else {
          upd[j] *= std::sqrt(temp);
        }
      }
    }
  }
  for (int j = 0; j < c; j++) perm[j] = j;
  for (int k = 0; k < size; k++) std::swap(perm[k], perm[transp[k]]);  // applyTranspositionOnTheRight(k, transp[k])
  return nonzero_pivots;
}

// HouseholderSequence::evalTo(dense): Q (r x r) = H_0 H_1 ... H_{nv-1}, reflectors applied to the
// identity from the last one backwards.  QR is the packed factor (r x c), nv = min(r, c).
inline Dense householder_q(const double* QR, int r, int c, const double* tau) {
  const int nv = std::min(r, c);
  Dense Q = Dense::identity(r, r);
  for (int k = nv - 1; k >= 0; --k) {
    const int corner = r - k;
    apply_householder_left(Q.v.data() + (size_t)k * r + k, r, corner, corner, QR + (size_t)k * r + k + 1, tau[k]);
  }
  return Q;
}

// y = Q^T y  with Q = H_0 ... H_{nv-1}  (apply H_0 first).
inline void apply_qt_inplace(const double* QR, int r, int c, const double* tau, double* y, int ldy, int nrhs) {
  const int nv = std::min(r, c);
  for (int k = 0; k < nv; k++) apply_householder_left(y + k, ldy, r - k, nrhs, QR + (size_t)k * r + k + 1, tau[k]);
}
// y = Q y (apply H_{nv-1} first).
inline void apply_q_inplace(const double* QR, int r, int c, const double* tau, double* y, int ldy, int nrhs) {
  const int nv = std::min(r, c);
  for (int k = nv - 1; k >= 0; --k)
    apply_householder_left(y + k, ldy, r - k, nrhs, QR + (size_t)k * r + k + 1, tau[k]);
}

// ----------------------------------------------------------------------------------------
// Sparse containers (Eigen::SparseMatrix in compressed mode)
// ----------------------------------------------------------------------------------------
struct Sparse {            // row_major == true: CSR (outer = rows); false: CSC (outer = cols)
  bool row_major = false;
  int rows = 0, cols = 0;
  std::vector<int> outer;  // size outerSize+1
  std::vector<int> inner;
  std::vector<double> val;
  int outer_size() const { return row_major ? rows : cols; }
  size_t nnz() const { return val.size(); }
};

struct Triplet { int r, c; double v; };

// Eigen::SparseMatrix::setFromTriplets: compressed result, inner indices sorted within each outer
// vector, duplicates summed (explicit zeros are kept).
inline Sparse from_triplets(int rows, int cols, const std::vector<Triplet>& t, bool row_major) {
  Sparse S;
  S.row_major = row_major; S.rows = rows; S.cols = cols;
  const int no = S.outer_size();
  std::vector<std::vector<std::pair<int, double>>> buckets(no);
  for (const auto& e : t) {
    const int o = row_major ? e.r : e.c, in = row_major ? e.c : e.r;
    buckets[o].push_back({in, e.v});
  }
  S.outer.assign(no + 1, 0);
  for (int o = 0; o < no; o++) {
    auto& b = buckets[o];
    std::stable_sort(b.begin(), b.end(), [](const auto& a, const auto& c) { return a.first < c.first; });
    size_t i = 0;
    while (i < b.size()) {
      double s = b[i].second; size_t j = i + 1;
      while (j < b.size() && b[j].first == b[i].first) { s += b[j].second; j++; }
      S.inner.push_back(b[i].first); S.val.push_back(s);
      i = j;
    }
    S.outer[o + 1] = (int)S.inner.size();
  }
  return S;
}

// y = S^T * x for a row-major S (sparse Q^T * dense), nrhs columns, col-major x (ldx) and y (ldy).
inline void spmv_transposed(const Sparse& S, const double* x, int ldx, double* y, int ldy, int nrhs) {
  for (int j = 0; j < nrhs; j++) {
    double* yj = y + (size_t)j * ldy;
    for (int i = 0; i < S.cols; i++) yj[i] = 0.0;
    for (int r = 0; r < S.rows; r++) {
      const double xr = x[(size_t)j * ldx + r];
      for (int p = S.outer[r]; p < S.outer[r + 1]; p++) yj[S.inner[p]] += S.val[p] * xr;
    }
  }
}
inline void spmv(const Sparse& S, const double* x, int ldx, double* y, int ldy, int nrhs) {  // row-major S
  for (int j = 0; j < nrhs; j++)
    for (int r = 0; r < S.rows; r++) {
      double s = 0.0;
      for (int p = S.outer[r]; p < S.outer[r + 1]; p++) s += S.val[p] * x[(size_t)j * ldx + S.inner[p]];
      y[(size_t)j * ldy + r] = s;
    }
}

// R.topLeftCorner(n, n).triangularView<Upper>().solve(b) for a col-major sparse R (sorted inner).
inline void sparse_upper_solve(const Sparse& R, int n, double* b) {
  for (int j = n - 1; j >= 0; --j) {
    // locate the diagonal in column j
    double diag = 0.0;
    for (int p = R.outer[j]; p < R.outer[j + 1]; p++)
      if (R.inner[p] == j) diag = R.val[p];
    b[j] /= diag;
    const double bj = b[j];
    for (int p = R.outer[j]; p < R.outer[j + 1]; p++) {
      const int i = R.inner[p];
      if (i < j) b[i] -= R.val[p] * bj;
    }
  }
}

// ----------------------------------------------------------------------------------------
// Block diagonal  (BlockDiagonalSparseQR.h)
// ----------------------------------------------------------------------------------------
enum QFormat { FullQ = 0, BlockDiagonalQ = 1 };
enum Info { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };  // Eigen::ComputationInfo

// Block storage = the device "block-COO" layout: flat col-major values, per-block rows/cols/offsets.
struct BlockDiag {
  int nRows = 0, nCols = 0;           // SparseBlockDiagonal::rows()/cols()
  std::vector<int> br, bc;            // per-block sizes
  std::vector<int64_t> off;           // offset of block i in values
  std::vector<double> values;
  int size() const { return (int)br.size(); }
};

struct BlockDiagQR {
  Sparse Q;                 // row-major (MatrixQType)
  Sparse R;                 // col-major (MatrixRType)
  std::vector<int> colPerm; // m_outputPerm_c.indices()
  std::vector<int> rowPerm; // m_rowPerm.indices()
  int rank = 0;
  int info = Success;
  // compact per-block factors kept for the compact CPU-baseline variant and for GPU comparisons
  std::vector<double> packed;  // same layout as BlockDiag::values: R in the upper triangle, essentials below
  std::vector<double> tau;     // concatenated, block i at sum_{l<i} c_l
};

// BlockDiagonalSparseQR::factorize (BlockDiagonalSparseQR.h:415-547), reference-faithful:
// per block dense QR (pivoted or not) -> explicit r x r Q_i -> scatter into global sparse Q and R.
inline BlockDiagQR block_diagonal_factorize(const BlockDiag& mat, bool colpiv, int qformat, bool build_q = true) {
  BlockDiagQR out;
  const int nb = mat.size();
  out.colPerm.resize(mat.nCols);
  std::iota(out.colPerm.begin(), out.colPerm.end(), 0);      // :417
  out.rowPerm.resize(mat.nRows);
  std::iota(out.rowPerm.begin(), out.rowPerm.end(), 0);      // analyzePattern :394-400 (no perm given)
  out.packed = mat.values;
  out.tau.assign(mat.nCols, 0.0);
  std::vector<Triplet> tripR;
  Sparse& Q = out.Q;
  Q.row_major = true; Q.rows = mat.nRows; Q.cols = mat.nRows;
  Q.outer.assign(mat.nRows + 1, 0);
  int m1 = 0;
  const int N_start = mat.nCols;
  int base_row = 0, base_col = 0, rank = 0;
  for (int i = 0; i < nb; i++) {
    const int r = mat.br[i], c = mat.bc[i];
    if (r < c) { out.info = InvalidInput; return out; }      // :509-516 landscape blocks rejected
    if (qformat != FullQ && qformat != BlockDiagonalQ) { out.info = InvalidInput; return out; }
    double* A = out.packed.data() + mat.off[i];
    double* tau = out.tau.data() + base_col;
    std::vector<int> p(c);
    if (colpiv) colpiv_householder_qr(A, r, c, tau, p.data());
    else { householder_qr(A, r, c, tau); std::iota(p.begin(), p.end(), 0); }
    rank += c;                                                // :440
    if (build_q) {
      Dense Qi = householder_q(A, r, c, tau);                 // :446
      const int curr_m1 = r - c;
      for (int j = 0; j < r; j++) {                           // :457-470 / :483-491
        if (qformat == FullQ) {
          for (int k = 0; k < c; k++) { Q.inner.push_back(base_col + k); Q.val.push_back(Qi(j, k)); }
          for (int k = 0; k < curr_m1; k++) { Q.inner.push_back(N_start + m1 + k); Q.val.push_back(Qi(j, c + k)); }
        } else {
          for (int k = 0; k < r; k++) { Q.inner.push_back(base_row + k); Q.val.push_back(Qi(j, k)); }
        }
        Q.outer[base_row + j + 1] = (int)Q.inner.size();
      }
      m1 += curr_m1;
    }
    for (int j = 0; j < c; j++)                               // :475-479 / :496-500
      for (int k = j; k < c; k++)
        tripR.push_back({(qformat == FullQ ? base_col : base_row) + j, base_col + k, A[(size_t)k * r + j]});
    for (int j = 0; j < c; j++) out.colPerm[base_col + j] = base_col + p[j];   // :519-521
    base_row += r; base_col += c;
  }
  if (build_q)
    for (int i = base_row; i < mat.nRows; i++) {              // :530-533 identity tail
      Q.inner.push_back(i); Q.val.push_back(1.0);
      Q.outer[i + 1] = (int)Q.inner.size();
    }
  out.R = from_triplets(mat.nRows, mat.nCols, tripR, false);  // :538-541
  out.rank = rank;
  out.info = Success;
  return out;
}

// BlockDiagonalSparseQR::_solve_impl (:258-280): y = Q^T b; y[0:rank] = R^-1 y[0:rank]; x = P_c * y[0:cols].
inline std::vector<double> block_diagonal_solve(const BlockDiagQR& f, int rows, int cols, const double* b) {
  std::vector<double> y(std::max(rows, cols), 0.0);
  spmv_transposed(f.Q, b, rows, y.data(), (int)y.size(), 1);
  sparse_upper_solve(f.R, f.rank, y.data());
  for (size_t i = f.rank; i < y.size(); i++) y[i] = 0.0;
  std::vector<double> x(cols);
  for (int j = 0; j < cols; j++) x[f.colPerm[j]] = y[j];     // dest = P * y  =>  dest(p[j]) = y(j)
  return x;
}

// Compact variant (same arithmetic per block, no explicit Q / sparse assembly): what the GPU computes.
// Writes packed factors + tau (+perm) in place and x.  FullQ y layout also produced when y != nullptr.
inline void block_diagonal_compact_factor_solve(BlockDiag& mat, bool colpiv, const double* b, double* x,
                                                double* tau_out, int* perm_out, double* y_full) {
  const int nb = mat.size();
  int base_row = 0, base_col = 0, m1 = 0;
  std::vector<double> bb;
  std::vector<int> p;
  for (int i = 0; i < nb; i++) {
    const int r = mat.br[i], c = mat.bc[i];
    double* A = mat.values.data() + mat.off[i];
    double* tau = tau_out + base_col;
    p.resize(c);
    if (colpiv) colpiv_householder_qr(A, r, c, tau, p.data());
    else { householder_qr(A, r, c, tau); std::iota(p.begin(), p.end(), 0); }
    if (perm_out) for (int j = 0; j < c; j++) perm_out[base_col + j] = base_col + p[j];
    if (b) {
      bb.assign(b + base_row, b + base_row + r);
      apply_qt_inplace(A, r, c, tau, bb.data(), r, 1);
      if (y_full) {
        for (int k = 0; k < c; k++) y_full[base_col + k] = bb[k];
        for (int k = 0; k < r - c; k++) y_full[mat.nCols + m1 + k] = bb[c + k];
      }
      for (int j = c - 1; j >= 0; --j) {
        double s = bb[j];
        for (int k = j + 1; k < c; k++) s -= A[(size_t)k * r + j] * bb[k];
        bb[j] = s / A[(size_t)j * r + j];
      }
      for (int j = 0; j < c; j++) x[base_col + p[j]] = bb[j];
    }
    base_row += r; base_col += c; m1 += r - c;
  }
}

// ----------------------------------------------------------------------------------------
// Block structure for banded matrices (SparseQRUtils.h:274-385)
// ----------------------------------------------------------------------------------------
struct BlockInfo { int idxRow = 0, idxCol = 0, numRows = 0, numCols = 0; };

inline void merge_blocks(std::vector<int>& order, std::map<int, BlockInfo>& bmap, int maxColStep, int suggestedBlockCols) {
  std::map<int, BlockInfo> nmap;
  std::vector<int> norder;
  BlockInfo first;
  int currRows = 0, currCols = 0;
  for (size_t it = 0; it < order.size(); ++it) {
    BlockInfo curr = bmap.at(order[it]);
    if (!norder.empty()) {
      BlockInfo last = nmap[norder.back()];
      if (curr.idxCol + curr.numCols <= last.idxCol + last.numCols) {
        nmap[norder.back()] = BlockInfo{last.idxRow, last.idxCol, last.numRows + curr.numRows, last.numCols};
        continue;
      }
    }
    if (first.numRows == 0) {
      first = curr; currRows = curr.numRows; currCols = curr.numCols;
    } else {
      currRows = curr.idxRow + curr.numRows - first.idxRow;
      currCols = curr.idxCol + curr.numCols - first.idxCol;
    }
    if (currRows > currCols && currCols >= maxColStep && currCols >= suggestedBlockCols) {
      norder.push_back(first.idxCol);
      nmap.insert({first.idxCol, BlockInfo{first.idxRow, first.idxCol, currRows, currCols}});
      first = BlockInfo();
    }
  }
  if (first.numRows != 0) {
    if (currRows > currCols && currCols >= maxColStep && currCols >= suggestedBlockCols) {
      norder.push_back(first.idxCol);
      nmap.insert({first.idxCol, BlockInfo{first.idxRow, first.idxCol, currRows, currCols}});
    } else {
      BlockInfo last = nmap[norder.back()];
      nmap[norder.back()] = BlockInfo{last.idxRow, last.idxCol, last.numRows + currRows,
                                      first.idxCol + currCols - last.idxCol};
    }
  }
  order = norder; bmap = nmap;
}

// BlockBandedMatrixInfo::fromBlockBandedPattern (SparseQRUtils.h:274-302): returns the merged blocks in order.
inline std::vector<BlockInfo> from_block_banded_pattern(int matRows, int matCols, int blockRows, int blockCols,
                                                        int blockOverlap, int suggestedBlockCols) {
  (void)matRows;
  const int maxColStep = blockCols - blockOverlap;
  const int numBlocks = matCols / maxColStep;
  std::map<int, BlockInfo> bmap;
  std::vector<int> order;
  for (int i = 0; i < numBlocks; i++) {
    const int rowIdx = i * blockRows, colIdx = i * maxColStep;
    order.push_back(colIdx);
    bmap.insert({colIdx, BlockInfo{rowIdx, colIdx, blockRows, i < numBlocks - 1 ? blockCols : blockCols - blockOverlap}});
  }
  merge_blocks(order, bmap, maxColStep, suggestedBlockCols);
  std::vector<BlockInfo> res;
  for (int k : order) res.push_back(bmap.at(k));
  return res;
}

// BlockBandedMatrixInfo::fromBlockDiagonalPattern (SparseQRUtils.h:255-272)
inline std::vector<BlockInfo> from_block_diagonal_pattern(int matRows, int matCols, int blockRows, int blockCols) {
  (void)matRows;
  std::vector<BlockInfo> res;
  const int numBlocks = matCols / blockCols;
  for (int i = 0; i < numBlocks; i++) res.push_back(BlockInfo{i * blockRows, i * blockCols, blockRows, blockCols});
  return res;
}

// ----------------------------------------------------------------------------------------
// Compact-WY block storage (BlockYTY.h / SparseBlockYTY.h)
// ----------------------------------------------------------------------------------------
struct YTYBlock {
  Dense Y, T;      // Y: rows x cols unit lower trapezoidal; T: cols x cols upper, T = -T_eigen so Q = I + Y T Y^T
  int row = 0, col = 0, numZeros = 0;
};

// SparseBlockYTY_VecProduct::evalTo (SparseBlockYTY.h:102-139) on one vector, in place.
inline void ytysequence_apply(const std::vector<YTYBlock>& blocks, double* v, bool transpose) {
  std::vector<double> seg, t1, t2;
  auto step = [&](const YTYBlock& B) {
    const int rows = B.Y.rows, cols = B.Y.cols;
    const int s0 = B.row, l0 = cols, s1 = B.row + cols + B.numZeros, l1 = rows - cols;
    seg.resize(rows);
    for (int i = 0; i < l0; i++) seg[i] = v[s0 + i];
    for (int i = 0; i < l1; i++) seg[l0 + i] = v[s1 + i];
    t1.assign(cols, 0.0); t2.assign(cols, 0.0);
    for (int k = 0; k < cols; k++) { double s = 0; for (int i = 0; i < rows; i++) s += B.Y(i, k) * seg[i]; t1[k] = s; }
    for (int k = 0; k < cols; k++) {
      double s = 0;
      for (int l = 0; l < cols; l++) s += (transpose ? B.T(l, k) : B.T(k, l)) * t1[l];
      t2[k] = s;
    }
    for (int i = 0; i < rows; i++) { double s = 0; for (int k = 0; k < cols; k++) s += B.Y(i, k) * t2[k]; seg[i] += s; }
    for (int i = 0; i < l0; i++) v[s0 + i] = seg[i];
    for (int i = 0; i < l1; i++) v[s1 + i] = seg[l0 + i];
  };
  if (transpose) for (size_t k = 0; k < blocks.size(); k++) step(blocks[k]);
  else for (size_t k = blocks.size(); k-- > 0;) step(blocks[k]);
}

// ----------------------------------------------------------------------------------------
// Banded blocked  (BandedBlockedSparseQR.h:443-519).  The (row-permuted) input is given as a
// column-major sparse matrix; block(...) .toDense() extracts windows.
// ----------------------------------------------------------------------------------------
inline Dense sparse_block_to_dense(const Sparse& A /*CSC*/, int r0, int c0, int nr, int nc) {
  Dense D(nr, nc);
  for (int j = 0; j < nc; j++) {
    const int cj = c0 + j;
    if (cj < 0 || cj >= A.cols) continue;
    for (int p = A.outer[cj]; p < A.outer[cj + 1]; p++) {
      const int i = A.inner[p] - r0;
      if (i >= 0 && i < nr) D(i, j) = A.val[p];
    }
  }
  return D;
}

struct BandedQR {
  std::vector<YTYBlock> blocks;
  Sparse R;  // col-major, explicit zeros kept
  int rank = 0, info = Success;
  int rows = 0, cols = 0;
};

inline BandedQR banded_factorize(const Sparse& pmat /*CSC, rows already permuted*/, const std::vector<BlockInfo>& blocks) {
  BandedQR out;
  out.rows = pmat.rows; out.cols = pmat.cols;
  std::vector<Triplet> Rvals;
  const int numBlocks = (int)blocks.size();
  BlockInfo bi = blocks[0];
  Dense Ji = sparse_block_to_dense(pmat, bi.idxRow, bi.idxCol, bi.numRows, bi.numCols);   // :458
  int activeRows = bi.numRows, numZeros = 0;
  for (int i = 0; i < numBlocks; i++) {
    bi = blocks[i];
    std::vector<double> tau(std::min(Ji.rows, Ji.cols));
    householder_qr(Ji.v.data(), Ji.rows, Ji.cols, tau.data());                            // :468
    YTYBlock blk;
    blk.Y = Dense::identity(activeRows, bi.numCols);                                      // :471-475
    for (int bc = 0; bc < bi.numCols; bc++)
      for (int rr = bc + 1; rr < activeRows; rr++) blk.Y(rr, bc) = Ji(rr, bc);
    std::vector<double> tauc(bi.numCols, 0.0);
    for (int k = 0; k < (int)tau.size() && k < bi.numCols; k++) tauc[k] = tau[k];
    blk.T = block_householder_t_factor(blk.Y.v.data(), activeRows, activeRows, bi.numCols, tauc.data());
    for (auto& t : blk.T.v) t = -t;                                                       // :477
    const int diagIdx = bi.idxCol;
    blk.row = diagIdx; blk.col = diagIdx; blk.numZeros = numZeros;                        // :481
    out.blocks.push_back(blk);
    // V = upper triangular view of the packed factor (:484)
    const int solvedRows = (i == numBlocks - 1) ? bi.numRows : blocks[i + 1].idxCol - bi.idxCol;
    for (int br = 0; br < solvedRows; br++)
      for (int bc = 0; bc < bi.numCols; bc++)
        Rvals.push_back({diagIdx + br, bi.idxCol + bc, (br <= bc && br < Ji.rows) ? Ji(br, bc) : 0.0});
    if (i < numBlocks - 1) {                                                              // :494-507
      const BlockInfo biNext = blocks[i + 1];
      const int blockOverlap = (bi.idxCol + bi.numCols) - biNext.idxCol;
      const int colIncrement = bi.numCols - blockOverlap;
      activeRows = bi.numRows + biNext.numRows - colIncrement;
      numZeros = (biNext.idxRow + biNext.numRows) - activeRows - biNext.idxCol;
      numZeros = numZeros < 0 ? 0 : numZeros;
      const int numCols = (biNext.numCols >= blockOverlap) ? biNext.numCols : blockOverlap;
      Dense Jn = sparse_block_to_dense(pmat, bi.idxRow + colIncrement, biNext.idxCol, activeRows, numCols);
      if (blockOverlap > 0) {
        const int nr = activeRows - biNext.numRows;
        for (int c = 0; c < blockOverlap; c++)
          for (int r = 0; r < nr; r++)
            Jn(r, c) = (colIncrement + r <= colIncrement + c) ? Ji(colIncrement + r, colIncrement + c) : 0.0;
      }
      Ji = Jn;
    }
  }
  out.R = from_triplets(pmat.rows, pmat.cols, Rvals, false);                              // :511
  out.rank = pmat.cols;                                                                   // :514
  return out;
}

// BandedBlockedSparseQR::_solve_impl (:290-311) — b must already be row-permuted by the caller.
inline std::vector<double> banded_solve(const BandedQR& f, const double* b) {
  std::vector<double> y(b, b + f.rows);
  ytysequence_apply(f.blocks, y.data(), true);
  y.resize(std::max(f.rows, f.cols), 0.0);
  sparse_upper_solve(f.R, f.rank, y.data());
  return std::vector<double>(y.begin(), y.begin() + f.cols);
}

// ----------------------------------------------------------------------------------------
// Dense right-hand solvers for the block-angular border
// ----------------------------------------------------------------------------------------
struct DenseQR {           // ColPivHouseholderQR<MatrixXd> or BlockedThinDenseQR<MatrixXd, panel>
  int rows = 0, cols = 0;
  Dense packed;            // ColPiv: packed factor.  Thin: m_R (the updated matrix, R in the upper triangle)
  std::vector<double> tau; // ColPiv only
  std::vector<int> perm;   // colsPermutation().indices()
  std::vector<YTYBlock> blocks;  // Thin only
  bool thin = false;
  int rank = 0;
};

inline DenseQR dense_colpiv_compute(const Dense& A) {
  DenseQR q; q.rows = A.rows; q.cols = A.cols; q.packed = A;
  q.tau.assign(std::min(A.rows, A.cols), 0.0); q.perm.resize(A.cols);
  q.rank = colpiv_householder_qr(q.packed.v.data(), A.rows, A.cols, q.tau.data(), q.perm.data());
  return q;
}

// BlockedThinDenseQR::compute (BlockedThinDenseQR.h:104-176) with panel width `panel`.
inline DenseQR dense_blocked_thin_compute(const Dense& A, int panel) {
  DenseQR q; q.rows = A.rows; q.cols = A.cols; q.packed = A; q.thin = true;
  q.perm.resize(A.cols); std::iota(q.perm.begin(), q.perm.end(), 0);
  int solved = 0;
  Dense& M = q.packed;
  while (solved < M.cols) {
    int newCols = panel;                                       // updateBlockInfo :144-155
    int numRows = M.rows - solved;
    if (solved + newCols >= M.cols) newCols = M.cols - solved;
    // factorize :157-176 — HouseholderQR of the (numRows x newCols) panel at (solved, solved)
    Dense P = M.block(solved, solved, numRows, newCols);
    std::vector<double> tau(std::min(numRows, newCols));
    householder_qr(P.v.data(), numRows, newCols, tau.data());
    YTYBlock blk;
    blk.Y = Dense::identity(numRows, newCols);
    for (int bc = 0; bc < newCols; bc++) for (int rr = bc + 1; rr < numRows; rr++) blk.Y(rr, bc) = P(rr, bc);
    std::vector<double> tauc(newCols, 0.0);
    for (size_t k = 0; k < tau.size(); k++) tauc[k] = tau[k];
    blk.T = block_householder_t_factor(blk.Y.v.data(), numRows, numRows, newCols, tauc.data());
    for (auto& t : blk.T.v) t = -t;
    blk.row = solved; blk.col = solved; blk.numZeros = 0;
    // updateMat(idxRow, mat.cols(), ...) (BlockedThinQRBase.h:308-319): every column j >= solved,
    // col += Y (T^T (Y^T col))
    std::vector<double> t1(newCols), t2(newCols);
    for (int j = solved; j < M.cols; j++) {
      double* cj = M.col(j) + solved;
      for (int k = 0; k < newCols; k++) { double s = 0; for (int i = 0; i < numRows; i++) s += blk.Y(i, k) * cj[i]; t1[k] = s; }
      for (int k = 0; k < newCols; k++) { double s = 0; for (int l = 0; l < newCols; l++) s += blk.T(l, k) * t1[l]; t2[k] = s; }
      for (int i = 0; i < numRows; i++) { double s = 0; for (int k = 0; k < newCols; k++) s += blk.Y(i, k) * t2[k]; cj[i] += s; }
    }
    q.blocks.push_back(blk);
    solved += newCols;
  }
  q.rank = M.cols;
  return q;
}

// BlockedThinSparseQR::compute (BlockedThinSparseQR.h:105-166) with updateBlockInfo (:198-236) and factorize (:238-283), on a
// DENSE input (the border of the reference's test 6, test/test-qrkit.cpp:329-362, is fully dense):
//  * analyzePattern (:168-196): ColumnDensity and AsBandedAsPossible are stable sorts by stored entries per column / first
//    stored column per row; a fully dense matrix keeps both orders, so m_outputPerm_c and m_rowPerm start as the identity;
//  * updateBlockInfo: every column's last stored row is rows-1, so a panel always spans rows [m_nonzeroPivots, rows);
//  * factorize: ColPivHouseholderQR of the panel, its Y / T (computeBlockedRepresentation, BlockedThinQRBase.h:321-333),
//    updateMat on EVERY column from the panel's first one (:308-319), R column m_nonzeroPivots + bc = rows above the diagonal
//    position from the updated matrix (pivoted column) + the panel's triangle (:270-279), nonzero pivots appended to
//    m_nnzColPermIdxs, the others deferred to m_zeroColPermIdxs (:250-255), final P = [nonzero pivots ; deferred] (:150-158).
// Rank-deficient panels: the reference restarts R columns it has already written (startVec(m_nonzeroPivots + bc) collides with
// the deferred columns of the previous panel) and never applies later panels to a deferred column, i.e. its R is not a
// factor of A P there.  This restatement keeps everything the reference defines (pivot order, nonzero-pivot rule, deferral
// order, rank = m_nonzeroPivots) and gives the deferred columns what A P = Q R requires: every later block reflector is applied
// to them and they occupy the trailing columns of R.  With full column rank the two coincide entry by entry.
inline DenseQR dense_blocked_thin_sparse_compute(const Dense& A, int panel) {
  DenseQR q; q.rows = A.rows; q.cols = A.cols; q.thin = true;
  Dense M = A;                                                  // m_pmatDense
  std::vector<int> nnzIdx, zeroIdx;
  std::vector<int> zeroBlock;                                   // number of YTY blocks already applied to a deferred column
  Dense R(A.rows, A.cols);                                      // full columns of Q^T A P, position order
  int nzp = 0, solved = 0;
  while (solved < M.cols) {
    int newCols = panel;                                        // updateBlockInfo :198-236
    if (solved + newCols >= M.cols) newCols = M.cols - solved;
    const int idxRow = nzp, numRows = M.rows - nzp;
    Dense P = M.block(idxRow, solved, numRows, newCols);        // factorize :238-283
    const int size = std::min(numRows, newCols);
    std::vector<double> tau(std::max(size, 1), 0.0);
    std::vector<int> perm(newCols);
    const int nzp_p = colpiv_householder_qr(P.v.data(), numRows, newCols, tau.data(), perm.data());
    for (int c = 0; c < nzp_p; c++) nnzIdx.push_back(solved + perm[c]);
    for (int c = nzp_p; c < newCols; c++) { zeroIdx.push_back(solved + perm[c]); zeroBlock.push_back((int)q.blocks.size() + 1); }
    YTYBlock blk;
    blk.Y = Dense::identity(numRows, newCols);
    for (int bc = 0; bc < newCols; bc++) for (int rr = bc + 1; rr < numRows; rr++) blk.Y(rr, bc) = P(rr, bc);
    std::vector<double> tauc(newCols, 0.0);
    for (int k = 0; k < size; k++) tauc[k] = tau[k];
    blk.T = block_householder_t_factor(blk.Y.v.data(), numRows, numRows, newCols, tauc.data());
    for (auto& t : blk.T.v) t = -t;
    blk.row = idxRow; blk.col = solved; blk.numZeros = 0;
    std::vector<double> t1(newCols), t2(newCols);
    for (int j = solved; j < M.cols; j++) {                     // updateMat :308-319
      double* cj = M.col(j) + idxRow;
      for (int k = 0; k < newCols; k++) { double s = 0; for (int i = 0; i < numRows; i++) s += blk.Y(i, k) * cj[i]; t1[k] = s; }
      for (int k = 0; k < newCols; k++) { double s = 0; for (int l = 0; l < newCols; l++) s += blk.T(l, k) * t1[l]; t2[k] = s; }
      for (int i = 0; i < numRows; i++) { double s = 0; for (int k = 0; k < newCols; k++) s += blk.Y(i, k) * t2[k]; cj[i] += s; }
    }
    q.blocks.push_back(blk);
    for (int bc = 0; bc < nzp_p; bc++) {                        // R columns of the nonzero pivots (:270-279)
      const int pos = nzp + bc;
      for (int br = 0; br < nzp; br++) R(br, pos) = M(br, solved + perm[bc]);
      for (int br = 0; br <= bc; br++) R(nzp + br, pos) = P(br, bc);
    }
    nzp += nzp_p;
    solved += newCols;
  }
  for (size_t z = 0; z < zeroIdx.size(); z++) {                 // deferred columns: the remaining reflectors, then the tail of R
    std::vector<double> col(M.col(zeroIdx[z]), M.col(zeroIdx[z]) + M.rows);
    std::vector<YTYBlock> rest(q.blocks.begin() + zeroBlock[z], q.blocks.end());
    ytysequence_apply(rest, col.data(), true);
    for (int r = 0; r < M.rows; r++) R(r, nzp + (int)z) = col[r];
  }
  q.perm = nnzIdx;
  q.perm.insert(q.perm.end(), zeroIdx.begin(), zeroIdx.end());
  q.packed = R;
  q.rank = nzp;                                                 // rank() = m_nonzeroPivots
  return q;
}

inline void dense_qr_apply_qt(const DenseQR& q, double* v) {
  if (q.thin) ytysequence_apply(q.blocks, v, true);
  else apply_qt_inplace(q.packed.v.data(), q.rows, q.cols, q.tau.data(), v, q.rows, 1);
}

// ----------------------------------------------------------------------------------------
// Block angular  (BlockAngularSparseQR.h:459-514, dense border)
//   left solver = block diagonal (FullQ) or banded; right solver = dense ColPiv or blocked thin.
// ----------------------------------------------------------------------------------------
struct AngularQR {
  int n = 0, m1 = 0, m2 = 0, n1 = 0;
  bool left_banded = false;
  BlockDiagQR leftBD;
  BandedQR leftBanded;
  DenseQR right;
  Dense J2;                 // Q1^T J2 (n x m2)
  Sparse R;                 // col-major
  std::vector<int> colPerm, rowPerm;
  int rank = 0;
};

inline void angular_left_apply_qt(const AngularQR& f, double* v /*n1*/) {
  if (f.left_banded) ytysequence_apply(f.leftBanded.blocks, v, true);
  else {
    std::vector<double> y(f.n1);
    spmv_transposed(f.leftBD.Q, v, f.n1, y.data(), f.n1, 1);
    std::copy(y.begin(), y.end(), v);
  }
}

inline void angular_finish(AngularQR& f, const Dense& J2in, int right_kind, int panel) {
  const int n = f.n, m1 = f.m1, m2 = f.m2;
  f.J2 = J2in;                                               // solveRightBlock :361-369 (left row perm = identity)
  for (int j = 0; j < m2; j++) angular_left_apply_qt(f, f.J2.col(j));
  Dense Abot = f.J2.block(m1, 0, n - m1, m2);
  f.right = right_kind == 0 ? dense_colpiv_compute(Abot) : right_kind == 2 ? dense_blocked_thin_sparse_compute(Abot, panel) : dense_blocked_thin_compute(Abot, panel);
  const Sparse& R1 = f.left_banded ? f.leftBanded.R : f.leftBD.R;
  Sparse& R = f.R;                                            // makeR :285-308
  R.row_major = false; R.rows = n; R.cols = m1 + m2; R.outer.assign(m1 + m2 + 1, 0);
  for (int c = 0; c < m1; c++) {
    for (int p = R1.outer[c]; p < R1.outer[c + 1]; p++) { R.inner.push_back(R1.inner[p]); R.val.push_back(R1.val[p]); }
    R.outer[c + 1] = (int)R.inner.size();
  }
  for (int c = 0; c < m2; c++) {
    for (int r = 0; r < m1; r++) { R.inner.push_back(r); R.val.push_back(f.J2(r, f.right.perm[c])); }
    for (int r = m1; r <= m1 + c; r++) { R.inner.push_back(r); R.val.push_back(f.right.packed(r - m1, c)); }
    R.outer[m1 + c + 1] = (int)R.inner.size();
  }
  f.colPerm.resize(m1 + m2);                                  // :498-503
  const std::vector<int>& p1 = f.left_banded ? std::vector<int>() : f.leftBD.colPerm;
  for (int j = 0; j < m1; j++) f.colPerm[j] = f.left_banded ? j : p1[j];
  for (int j = 0; j < m2; j++) f.colPerm[m1 + j] = m1 + f.right.perm[j];
  f.rowPerm.resize(n); std::iota(f.rowPerm.begin(), f.rowPerm.end(), 0);
  f.rank = (f.left_banded ? f.leftBanded.rank : f.leftBD.rank) + f.right.rank;   // :510
}

inline AngularQR angular_factorize_bd(const BlockDiag& J1, const Dense& J2, bool left_colpiv, int right_kind, int panel) {
  AngularQR f; f.n = J1.nRows; f.n1 = J1.nRows; f.m1 = J1.nCols; f.m2 = J2.cols; f.left_banded = false;
  f.leftBD = block_diagonal_factorize(J1, left_colpiv, FullQ);
  angular_finish(f, J2, right_kind, panel);
  return f;
}
inline AngularQR angular_factorize_banded(const Sparse& J1, const std::vector<BlockInfo>& blocks, const Dense& J2,
                                          int right_kind, int panel) {
  AngularQR f; f.n = J1.rows; f.n1 = J1.rows; f.m1 = J1.cols; f.m2 = J2.cols; f.left_banded = true;
  f.leftBanded = banded_factorize(J1, blocks);
  angular_finish(f, J2, right_kind, panel);
  return f;
}

// Q^T v (:607-625) then the bordered triangular solve and column permutation (:203-227).
inline std::vector<double> angular_apply_qt(const AngularQR& f, const double* v) {
  std::vector<double> res(v, v + f.n);
  angular_left_apply_qt(f, res.data());
  dense_qr_apply_qt(f.right, res.data() + f.m1);
  return res;
}
inline std::vector<double> angular_solve(const AngularQR& f, const double* b) {
  std::vector<double> y = angular_apply_qt(f, b);
  const int cols = f.m1 + f.m2;
  y.resize(std::max(f.n, cols), 0.0);
  sparse_upper_solve(f.R, f.rank, y.data());
  std::fill(y.begin() + f.rank, y.end(), 0.0);              // y.bottomRows(y.rows() - rank).setZero() (:216)
  std::vector<double> x(cols);
  for (int j = 0; j < cols; j++) x[f.colPerm[j]] = y[j];
  return x;
}

// ----------------------------------------------------------------------------------------
// Shared counter-based input generator: U[0.5, 5.0) as the reference tests
// (test/test-qrkit.cpp:65).
// ----------------------------------------------------------------------------------------
inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline double synth_value(uint64_t seed, uint64_t block, uint64_t row, uint64_t col, double lo = 0.5, double hi = 5.0) {
  const uint64_t u = splitmix64(seed ^ (block << 20) ^ (row << 10) ^ col);
  return lo + (hi - lo) * (double)(u >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace qrk_oracle
