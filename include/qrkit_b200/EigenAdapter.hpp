// EigenAdapter.hpp — the reference-side binding of qrkit_b200: Eigen-typed solver classes with the template signatures and
// members of jasvob/QRKit's own solvers, forwarding to the C ABI (include/qrkit_b200.h).  This is the file a QRKit
// maintainer drops next to src/QRKit/*.h (INTEGRATION.md walks through it).
//
//   QRKit::BlockDiagonalSparseQR_B200<BlockQRSolver, QFormat>   <->  BlockDiagonalSparseQR   (BlockDiagonalSparseQR.h:37-335)
//   QRKit::BlockAngularSparseQR_B200<LeftSolver, RightSolver>   <->  BlockAngularSparseQR    (BlockAngularSparseQR.h:79-281)
//   QRKit::BandedBlockedSparseQR_B200<BlockQRSolver, Overlap>   <->  BandedBlockedSparseQR   (BandedBlockedSparseQR.h:122-344)
//
// All three derive from Eigen::SparseSolverBase (CRTP, `BlockDiagonalSparseQR.h:38`), provide `_solve_impl` so that
// `solver.solve(b)` returns an Eigen::Solve expression (:258-299), and specialise SparseQRUtils::HasRowsPermutation
// (:337-340) so that they can serve as LeftSolver of the reference's own BlockAngularSparseQR.
//
// Needs Eigen (>= 3.3) and the reference's SparseBlockDiagonal.h / BlockMatrix1x2.h / SparseQRUtils.h on the include path;
// it is inert otherwise.  Eigen is not part of this repository's image: tests/test_eigen_adapter_cpu.py compiles this header
// against a minimal mock of the Eigen / QRKit names it touches (tests/cpp/mock_eigen) to keep it syntactically honest.
#ifndef QRKIT_B200_EIGEN_ADAPTER_HPP_
#define QRKIT_B200_EIGEN_ADAPTER_HPP_

#if defined(__has_include)
#if __has_include(<Eigen/Sparse>)
#define QRKIT_B200_HAVE_EIGEN 1
#endif
#endif

#ifdef QRKIT_B200_HAVE_EIGEN
#include <Eigen/Dense>
#include <Eigen/Sparse>

#include <cstdint>
#include <type_traits>
#include <vector>

#include "../qrkit_b200.h"

namespace QRKit {

namespace b200_detail {
// Eigen::HouseholderQR<...> -> no pivoting; anything else (ColPivHouseholderQR, the tests' wrappers) -> Eigen's ColPiv rule
template <typename Solver> struct Pivoting { enum { value = QRK_PIVOT_COLPIV }; };
template <typename M> struct Pivoting<Eigen::HouseholderQR<M> > { enum { value = QRK_PIVOT_NONE }; };

typedef Eigen::SparseMatrix<double, Eigen::ColMajor, int> SpColMajor;
typedef Eigen::SparseMatrix<double, Eigen::RowMajor, int> SpRowMajor;
typedef Eigen::PermutationMatrix<Eigen::Dynamic, Eigen::Dynamic, int> Perm;

inline Eigen::ComputationInfo to_info(int status, qrk_handle_t h) {
  if (status != QRK_STATUS_OK)      // no exceptions, as the reference: a failed call surfaces through info()
    return status == QRK_STATUS_INVALID_ARGUMENT || status == QRK_STATUS_UNSUPPORTED || status == QRK_STATUS_NO_DEVICE ? Eigen::InvalidInput : Eigen::NumericalIssue;
  int32_t i = QRK_INFO_SUCCESS;
  if (h) qrk_info(h, &i);
  return (Eigen::ComputationInfo)i;
}
// matrixR(): the library returns the reference's exact compressed layout; copy it into an owned Eigen matrix
inline SpColMajor fetch_r(qrk_handle_t h, Eigen::Index rows, Eigen::Index cols) {
  int64_t nnz = 0;
  qrk_matrix_r_nnz(h, &nnz);
  Eigen::VectorXi outer(cols + 1), inner(nnz > 0 ? nnz : 1);
  Eigen::VectorXd val(nnz > 0 ? nnz : 1);
  qrk_matrix_r(h, outer.data(), inner.data(), val.data(), QRK_HOST);
  SpColMajor R = Eigen::Map<const SpColMajor>(rows, cols, nnz, outer.data(), inner.data(), val.data());
  return R;
}
inline Perm fetch_perm(qrk_handle_t h, Eigen::Index n, bool cols) {
  Perm p(n);
  if (cols) qrk_cols_permutation(h, p.indices().data(), QRK_HOST);
  else qrk_rows_permutation(h, p.indices().data(), QRK_HOST);
  return p;
}
// matrixQ(): expression object; `.transpose() * dense` and `* dense` forward to qrk_apply_qt / qrk_apply_q
// (BlockDiagonalSparseQR.h:235-237 used at :266; BlockAngularSparseQR.h:651-701)
template <typename QR>
struct QExpr {
  const QR& qr;
  bool transposed;
  QExpr transpose() const { return QExpr{qr, !transposed}; }
  QExpr adjoint() const { return transpose(); }
  Eigen::Index rows() const { return qr.rows(); }
  Eigen::Index cols() const { return qr.rows(); }
  Eigen::MatrixXd operator*(const Eigen::MatrixXd& B) const {
    Eigen::MatrixXd Y(qr.rows(), B.cols());
    (transposed ? qrk_apply_qt : qrk_apply_q)(qr.handle(), B.data(), B.rows(), Y.data(), Y.rows(), (int32_t)B.cols(), QRK_HOST);
    return Y;
  }
};
}  // namespace b200_detail

// ---------------------------------------------------------------------------------------------------------------------
// BlockDiagonalSparseQR on the GPU.  Fixed-size blocks: SparseBlockDiagonal<Matrix<double,R,C>> keeps them in one
// std::vector, which IS the block-COO value array (SparseBlockDiagonal.h:46,159) — uploaded with one copy.
// ---------------------------------------------------------------------------------------------------------------------
template <typename _BlockQRSolver, int _QFormat = 0>
class BlockDiagonalSparseQR_B200 : public Eigen::SparseSolverBase<BlockDiagonalSparseQR_B200<_BlockQRSolver, _QFormat> > {
 protected:
  typedef Eigen::SparseSolverBase<BlockDiagonalSparseQR_B200<_BlockQRSolver, _QFormat> > Base;
  using Base::m_isInitialized;

 public:
  using Base::_solve_impl;
  typedef _BlockQRSolver BlockQRSolver;
  typedef typename BlockQRSolver::MatrixType BlockMatrixType;
  typedef SparseBlockDiagonal<BlockMatrixType> MatrixType;
  typedef double Scalar;
  typedef double RealScalar;
  typedef int StorageIndex;
  typedef Eigen::Index Index;
  typedef b200_detail::SpRowMajor MatrixQType;
  typedef b200_detail::SpColMajor MatrixRType;
  typedef b200_detail::Perm PermutationType;
  enum { ColsAtCompileTime = Eigen::Dynamic, MaxColsAtCompileTime = Eigen::Dynamic };

  BlockDiagonalSparseQR_B200() {}
  explicit BlockDiagonalSparseQR_B200(const MatrixType& mat) { compute(mat); }
  ~BlockDiagonalSparseQR_B200() { qrk_destroy(m_h); }
  BlockDiagonalSparseQR_B200(const BlockDiagonalSparseQR_B200&) = delete;
  BlockDiagonalSparseQR_B200& operator=(const BlockDiagonalSparseQR_B200&) = delete;

  void compute(const MatrixType& mat, const PermutationType& rowPerm = PermutationType(), bool /*forcePatternAnalysis*/ = false) {   // :94-102
    analyzePattern(mat, rowPerm);
    factorize(mat);
  }
  void analyzePattern(const MatrixType& mat, const PermutationType& rowPerm = PermutationType()) {                               // :392-405
    qrk_desc_t d = qrk_desc_t();
    d.kind = QRK_BLOCK_DIAGONAL;
    d.num_blocks = (int64_t)mat.size();
    d.block_rows = BlockMatrixType::RowsAtCompileTime;
    d.block_cols = BlockMatrixType::ColsAtCompileTime;
    d.n_rows = mat.rows();
    d.n_cols = mat.cols();
    d.q_format = _QFormat;
    d.pivoting = b200_detail::Pivoting<BlockQRSolver>::value;
    qrk_destroy(m_h);
    m_h = nullptr;
    m_info = b200_detail::to_info(qrk_create(&d, &m_h), nullptr);
    if (!m_h) return;
    qrk_analyze_pattern(m_h, rowPerm.size() ? rowPerm.indices().data() : nullptr);
    m_rows = mat.rows();
    m_cols = mat.cols();
  }
  void factorize(const MatrixType& mat) {                                                                                          // :415-547
    if (!m_h) return;
    int st = qrk_set_blocks(m_h, mat.size() ? mat[0].data() : nullptr, QRK_HOST);
    if (st == QRK_STATUS_OK) st = qrk_factorize(m_h);
    m_info = b200_detail::to_info(st, m_h);
    m_haveR = false;
    m_isInitialized = (st == QRK_STATUS_OK);
  }
  Index rows() const { return m_rows; }
  Index cols() const { return m_cols; }
  Index rank() const { int64_t r = 0; qrk_rank(m_h, &r); return (Index)r; }
  Eigen::ComputationInfo info() const { return m_info; }
  const MatrixRType& matrixR() const {                                                                                             // :156
    if (!m_haveR) { m_R = b200_detail::fetch_r(m_h, m_rows, m_cols); m_haveR = true; }
    return m_R;
  }
  b200_detail::QExpr<BlockDiagonalSparseQR_B200> matrixQ() const { return b200_detail::QExpr<BlockDiagonalSparseQR_B200>{*this, false}; }
  PermutationType colsPermutation() const { return b200_detail::fetch_perm(m_h, m_cols, true); }                                  // :242
  PermutationType rowsPermutation() const { return b200_detail::fetch_perm(m_h, m_rows, false); }                                 // :251
  template <typename Rhs, typename Dest>
  bool _solve_impl(const Eigen::MatrixBase<Rhs>& B, Eigen::MatrixBase<Dest>& dest) const {                                        // :258-280
    eigen_assert(m_isInitialized && "The factorization should be called first, use compute()");
    Eigen::MatrixXd b = B, x(m_cols, B.cols());
    const int st = qrk_solve(m_h, b.data(), b.rows(), x.data(), m_cols, (int32_t)b.cols(), QRK_HOST);
    dest = x;
    m_info = b200_detail::to_info(st, m_h);
    return st == QRK_STATUS_OK;
  }
  template <typename Rhs>
  inline const Eigen::Solve<BlockDiagonalSparseQR_B200, Rhs> solve(const Eigen::MatrixBase<Rhs>& B) const {                       // :287-292
    eigen_assert(m_isInitialized && "The factorization should be called first, use compute()");
    return Eigen::Solve<BlockDiagonalSparseQR_B200, Rhs>(*this, B.derived());
  }
  qrk_handle_t handle() const { return m_h; }

 protected:
  qrk_handle_t m_h = nullptr;
  Index m_rows = 0, m_cols = 0;
  mutable Eigen::ComputationInfo m_info = Eigen::Success;
  mutable bool m_haveR = false;
  mutable MatrixRType m_R;
};

template <typename _BlockQRSolver, int _QFormat>
struct SparseQRUtils::HasRowsPermutation<BlockDiagonalSparseQR_B200<_BlockQRSolver, _QFormat> > {
  static const bool value = true;
};

// ---------------------------------------------------------------------------------------------------------------------
// BlockAngularSparseQR on the GPU: A = [J1 | J2], J1 block diagonal with fixed-size blocks, J2 dense.
// _LeftSolver: a BlockDiagonalSparseQR(_B200)<BlockQRSolver> type (its BlockQRSolver fixes the block size and the pivoting);
// _RightSolver: Eigen::ColPivHouseholderQR<MatrixXd> (QRK_RIGHT_COLPIV), BlockedThinDenseQR / Eigen::HouseholderQR
// (QRK_RIGHT_UNPIVOTED), or BlockedThinSparseQR (QRK_RIGHT_THIN_SPARSE) — selected by the RightSolverKind trait below.
// ---------------------------------------------------------------------------------------------------------------------
template <typename RightSolver> struct RightSolverKind { enum { value = QRK_RIGHT_COLPIV }; };
template <typename M> struct RightSolverKind<Eigen::HouseholderQR<M> > { enum { value = QRK_RIGHT_UNPIVOTED }; };

template <typename _LeftSolver, typename _RightSolver>
class BlockAngularSparseQR_B200 : public Eigen::SparseSolverBase<BlockAngularSparseQR_B200<_LeftSolver, _RightSolver> > {
 protected:
  typedef Eigen::SparseSolverBase<BlockAngularSparseQR_B200<_LeftSolver, _RightSolver> > Base;
  using Base::m_isInitialized;

 public:
  using Base::_solve_impl;
  typedef _LeftSolver BlockQRSolverLeft;
  typedef _RightSolver BlockQRSolverRight;
  typedef typename BlockQRSolverLeft::MatrixType LeftBlockMatrixType;
  typedef Eigen::MatrixXd RightBlockMatrixType;
  typedef BlockMatrix1x2<LeftBlockMatrixType, RightBlockMatrixType> MatrixType;
  typedef typename BlockQRSolverLeft::BlockMatrixType BlockMatrixType;
  typedef double Scalar;
  typedef int StorageIndex;
  typedef Eigen::Index Index;
  typedef b200_detail::SpColMajor MatrixRType;
  typedef b200_detail::Perm PermutationType;
  enum { ColsAtCompileTime = Eigen::Dynamic, MaxColsAtCompileTime = Eigen::Dynamic };

  BlockAngularSparseQR_B200() {}
  explicit BlockAngularSparseQR_B200(const MatrixType& mat) { compute(mat); }
  ~BlockAngularSparseQR_B200() { qrk_destroy(m_h); }
  BlockAngularSparseQR_B200(const BlockAngularSparseQR_B200&) = delete;
  BlockAngularSparseQR_B200& operator=(const BlockAngularSparseQR_B200&) = delete;

  void compute(const MatrixType& mat) {                                                                                            // :134-138
    const LeftBlockMatrixType& L = mat.leftBlock();
    const RightBlockMatrixType& J2 = mat.rightBlock();
    qrk_desc_t d = qrk_desc_t();
    d.kind = QRK_BLOCK_ANGULAR;
    d.left_solver = QRK_LEFT_BLOCK_DIAGONAL;
    d.num_blocks = (int64_t)L.size();
    d.block_rows = BlockMatrixType::RowsAtCompileTime;
    d.block_cols = BlockMatrixType::ColsAtCompileTime;
    d.pivoting = b200_detail::Pivoting<typename BlockQRSolverLeft::BlockQRSolver>::value;
    d.q_format = QRK_FULL_Q;
    d.border_cols = (int32_t)J2.cols();
    d.right_solver = RightSolverKind<BlockQRSolverRight>::value;
    qrk_destroy(m_h);
    m_h = nullptr;
    int st = qrk_create(&d, &m_h);
    if (st == QRK_STATUS_OK) st = qrk_set_border(m_h, J2.data(), J2.rows(), QRK_HOST);
    if (st == QRK_STATUS_OK) st = qrk_compute(m_h, L.size() ? L[0].data() : nullptr, QRK_HOST);
    m_info = b200_detail::to_info(st, m_h);
    m_rows = mat.rows();
    m_cols = mat.cols();
    m_leftCols = L.cols();
    m_haveR = false;
    m_isInitialized = (st == QRK_STATUS_OK);
  }
  Index rows() const { return m_rows; }
  Index cols() const { return m_cols; }
  Index leftBlockCols() const { return m_leftCols; }                                                                              // :276-280
  Index rank() const { int64_t r = 0; qrk_rank(m_h, &r); return (Index)r; }                                                        // :510
  Eigen::ComputationInfo info() const { return m_info; }
  const MatrixRType& matrixR() const {                                                                                             // [R1, Atop P2; 0, R2] (:285-308)
    if (!m_haveR) { m_R = b200_detail::fetch_r(m_h, m_rows, m_cols); m_haveR = true; }
    return m_R;
  }
  b200_detail::QExpr<BlockAngularSparseQR_B200> matrixQ() const { return b200_detail::QExpr<BlockAngularSparseQR_B200>{*this, false}; }   // :598-644
  PermutationType colsPermutation() const { return b200_detail::fetch_perm(m_h, m_cols, true); }                                  // [P1; m1 + P2] (:498-503)
  PermutationType rowsPermutation() const { return b200_detail::fetch_perm(m_h, m_rows, false); }
  template <typename Rhs, typename Dest>
  bool _solve_impl(const Eigen::MatrixBase<Rhs>& B, Eigen::MatrixBase<Dest>& dest) const {                                        // :203-227
    eigen_assert(m_isInitialized && "The factorization should be called first, use compute()");
    Eigen::MatrixXd b = B, x(m_cols, B.cols());
    const int st = qrk_solve(m_h, b.data(), b.rows(), x.data(), m_cols, (int32_t)b.cols(), QRK_HOST);
    dest = x;
    m_info = b200_detail::to_info(st, m_h);
    return st == QRK_STATUS_OK;
  }
  template <typename Rhs>
  inline const Eigen::Solve<BlockAngularSparseQR_B200, Rhs> solve(const Eigen::MatrixBase<Rhs>& B) const {
    eigen_assert(m_isInitialized && "The factorization should be called first, use compute()");
    return Eigen::Solve<BlockAngularSparseQR_B200, Rhs>(*this, B.derived());
  }
  // One pass: factorize + Q^T b + both back substitutions (what an LM iteration asks for; no reference equivalent)
  Eigen::VectorXd computeAndSolve(const MatrixType& mat, const Eigen::VectorXd& b) {
    compute(mat);
    return solve(b);
  }
  qrk_handle_t handle() const { return m_h; }

 protected:
  qrk_handle_t m_h = nullptr;
  Index m_rows = 0, m_cols = 0, m_leftCols = 0;
  mutable Eigen::ComputationInfo m_info = Eigen::Success;
  mutable bool m_haveR = false;
  mutable MatrixRType m_R;
};

template <typename _LeftSolver, typename _RightSolver>
struct SparseQRUtils::HasRowsPermutation<BlockAngularSparseQR_B200<_LeftSolver, _RightSolver> > {
  static const bool value = true;
};

// ---------------------------------------------------------------------------------------------------------------------
// BandedBlockedSparseQR on the GPU for the block-banded pattern of fromBlockBandedPattern (SparseQRUtils.h:274-302):
// compute(const SparseMatrix&) slices the matrix into its block_rows x block_cols slabs (block row k at rows k*block_rows,
// columns k*(block_cols - overlap)) — the analyzePattern branch of BandedBlockedSparseQR.h:401-407.
// ---------------------------------------------------------------------------------------------------------------------
template <typename _MatrixType, typename _BlockQRSolver, int _BlockOverlap, int _SuggestedBlockCols = 2>
class BandedBlockedSparseQR_B200 : public Eigen::SparseSolverBase<BandedBlockedSparseQR_B200<_MatrixType, _BlockQRSolver, _BlockOverlap, _SuggestedBlockCols> > {
 protected:
  typedef Eigen::SparseSolverBase<BandedBlockedSparseQR_B200<_MatrixType, _BlockQRSolver, _BlockOverlap, _SuggestedBlockCols> > Base;
  using Base::m_isInitialized;

 public:
  using Base::_solve_impl;
  typedef _MatrixType MatrixType;
  typedef _BlockQRSolver BlockQRSolver;
  typedef double Scalar;
  typedef int StorageIndex;
  typedef Eigen::Index Index;
  typedef b200_detail::SpColMajor MatrixRType;
  typedef b200_detail::Perm PermutationType;
  enum { ColsAtCompileTime = Eigen::Dynamic, MaxColsAtCompileTime = Eigen::Dynamic,
         BlockRows = BlockQRSolver::RowsAtCompileTime, BlockCols = BlockQRSolver::ColsAtCompileTime };

  BandedBlockedSparseQR_B200() {}
  ~BandedBlockedSparseQR_B200() { qrk_destroy(m_h); }
  BandedBlockedSparseQR_B200(const BandedBlockedSparseQR_B200&) = delete;
  BandedBlockedSparseQR_B200& operator=(const BandedBlockedSparseQR_B200&) = delete;

  void compute(const MatrixType& mat) {                                                                                            // :170-178
    const Index step = BlockCols - _BlockOverlap;
    const Index nb = mat.rows() / BlockRows;
    std::vector<double> slabs((size_t)(nb * BlockRows * BlockCols), 0.0);
    for (Index k = 0; k < nb; k++)                                    // mat.block(k*BlockRows, k*step, BlockRows, BlockCols) per block row
      for (Index j = 0; j < BlockCols && k * step + j < mat.cols(); j++)
        for (typename MatrixType::InnerIterator it(mat, k * step + j); it; ++it) {
          const Index r = it.row() - k * BlockRows;
          if (r >= 0 && r < BlockRows) slabs[(size_t)((k * BlockCols + j) * BlockRows + r)] = it.value();
        }
    qrk_desc_t d = qrk_desc_t();
    d.kind = QRK_BANDED_BLOCKED;
    d.num_blocks = nb;
    d.block_rows = BlockRows;
    d.block_cols = BlockCols;
    d.block_overlap = _BlockOverlap;
    d.n_cols = mat.cols();                                             // the last slab may be narrower (SparseQRUtils.h:284)
    d.reserved[0] = _SuggestedBlockCols;
    qrk_destroy(m_h);
    m_h = nullptr;
    int st = qrk_create(&d, &m_h);
    if (st == QRK_STATUS_OK) st = qrk_compute(m_h, slabs.data(), QRK_HOST);
    m_info = b200_detail::to_info(st, m_h);
    m_rows = mat.rows();
    m_cols = mat.cols();
    m_haveR = false;
    m_isInitialized = (st == QRK_STATUS_OK);
  }
  Index rows() const { return m_rows; }
  Index cols() const { return m_cols; }
  Index rank() const { int64_t r = 0; qrk_rank(m_h, &r); return (Index)r; }                                                        // :514
  Eigen::ComputationInfo info() const { return m_info; }
  const MatrixRType& matrixR() const {                                                                                             // :484-491, explicit zeros included
    if (!m_haveR) { m_R = b200_detail::fetch_r(m_h, m_rows, m_cols); m_haveR = true; }
    return m_R;
  }
  PermutationType colsPermutation() const { return b200_detail::fetch_perm(m_h, m_cols, true); }
  PermutationType rowsPermutation() const { return b200_detail::fetch_perm(m_h, m_rows, false); }
  // (matrixQ().transpose() * b).topRows(cols()) and matrixQ() * [y; 0]: the thin factor, the Q products a banded factor has here
  Eigen::MatrixXd applyQtThin(const Eigen::MatrixXd& B) const {
    Eigen::MatrixXd Y(m_cols, B.cols());
    qrk_apply_qt_thin(m_h, B.data(), B.rows(), Y.data(), Y.rows(), (int32_t)B.cols(), QRK_HOST);
    return Y;
  }
  Eigen::MatrixXd applyQThin(const Eigen::MatrixXd& Y) const {
    Eigen::MatrixXd X(m_rows, Y.cols());
    qrk_apply_q_thin(m_h, Y.data(), Y.rows(), X.data(), X.rows(), (int32_t)Y.cols(), QRK_HOST);
    return X;
  }
  template <typename Rhs, typename Dest>
  bool _solve_impl(const Eigen::MatrixBase<Rhs>& B, Eigen::MatrixBase<Dest>& dest) const {                                        // :287-307
    eigen_assert(m_isInitialized && "The factorization should be called first, use compute()");
    Eigen::MatrixXd b = B, x(m_cols, B.cols());
    const int st = qrk_solve(m_h, b.data(), b.rows(), x.data(), m_cols, (int32_t)b.cols(), QRK_HOST);
    dest = x;
    m_info = b200_detail::to_info(st, m_h);
    return st == QRK_STATUS_OK;
  }
  template <typename Rhs>
  inline const Eigen::Solve<BandedBlockedSparseQR_B200, Rhs> solve(const Eigen::MatrixBase<Rhs>& B) const {
    eigen_assert(m_isInitialized && "The factorization should be called first, use compute()");
    return Eigen::Solve<BandedBlockedSparseQR_B200, Rhs>(*this, B.derived());
  }
  qrk_handle_t handle() const { return m_h; }

 protected:
  qrk_handle_t m_h = nullptr;
  Index m_rows = 0, m_cols = 0;
  mutable Eigen::ComputationInfo m_info = Eigen::Success;
  mutable bool m_haveR = false;
  mutable MatrixRType m_R;
};

template <typename _MatrixType, typename _BlockQRSolver, int _BlockOverlap, int _SuggestedBlockCols>
struct SparseQRUtils::HasRowsPermutation<BandedBlockedSparseQR_B200<_MatrixType, _BlockQRSolver, _BlockOverlap, _SuggestedBlockCols> > {
  static const bool value = true;
};

}  // namespace QRKit

#endif  // QRKIT_B200_HAVE_EIGEN
#endif  // QRKIT_B200_EIGEN_ADAPTER_HPP_
