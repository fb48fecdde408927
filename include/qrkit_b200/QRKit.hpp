// QRKit.hpp — header-only C++17 host façade over the C ABI (qrkit_b200.h) with the reference's solver
// interface: same class names, method names, argument meaning and error behaviour as
//   BlockDiagonalSparseQR  (reference src/QRKit/BlockDiagonalSparseQR.h:37-335)
//   BlockAngularSparseQR   (reference src/QRKit/BlockAngularSparseQR.h:79-281)
//   BandedBlockedSparseQR  (reference src/QRKit/BandedBlockedSparseQR.h:122-344), fixed block-banded pattern
//   SparseBlockDiagonal    (reference src/QRKit/SparseBlockDiagonal.h:44-163)
//   BlockMatrix1x2         (reference src/QRKit/BlockMatrix1x2.h:31-67)
// so that code written against QRKit's Eigen-style API (compute(), matrixQ().transpose() * b, matrixR(),
// solve(), rank(), info(), colsPermutation(), rowsPermutation()) switches by changing a namespace.
// Eigen is NOT required: the small value types below stand in for Eigen::VectorXd / MatrixXd /
// SparseMatrix / PermutationMatrix.  When <Eigen/Sparse> is available, include qrkit_b200/EigenAdapter.hpp
// for zero-copy maps between the two (untested in this image, which has no Eigen).
//
// All arithmetic happens in libqrkit_b200.so on the GPU; there is no CPU fallback: without a CUDA device
// every solver reports info() == InvalidInput and lastErrorMessage() says so.
#pragma once
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <exception>
#include <thread>
#include <vector>

#include "../qrkit_b200.h"

namespace QRKit_b200 {

// Eigen::ComputationInfo
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };

using Index = std::int64_t;
using StorageIndex = int;   // test/test-qrkit.cpp:40

// ---- value types standing in for Eigen's --------------------------------------------------------------
using VectorXd = std::vector<double>;

struct MatrixXd {           // column-major dense matrix
  Index m_rows = 0, m_cols = 0;
  std::vector<double> m_data;
  MatrixXd() {}
  MatrixXd(Index r, Index c) : m_rows(r), m_cols(c), m_data((size_t)r * c, 0.0) {}
  Index rows() const { return m_rows; }
  Index cols() const { return m_cols; }
  double& operator()(Index i, Index j) { return m_data[(size_t)j * m_rows + i]; }
  double operator()(Index i, Index j) const { return m_data[(size_t)j * m_rows + i]; }
  double* data() { return m_data.data(); }
  const double* data() const { return m_data.data(); }
};

template <int Rows, int Cols>
struct Matrix {              // Eigen::Matrix<double, Rows, Cols>: fixed-size column-major block, no padding
  enum { RowsAtCompileTime = Rows, ColsAtCompileTime = Cols };
  double m_data[Rows * Cols];
  static constexpr Index rows() { return Rows; }
  static constexpr Index cols() { return Cols; }
  double& operator()(Index i, Index j) { return m_data[j * Rows + i]; }
  double operator()(Index i, Index j) const { return m_data[j * Rows + i]; }
  const double* data() const { return m_data; }
};

template <int Major>        // 0 = ColMajor, 1 = RowMajor (Eigen's option values)
struct SparseMatrix {       // compressed storage, int32 indices
  Index m_rows = 0, m_cols = 0;
  std::vector<StorageIndex> outer, inner;
  std::vector<double> values;
  Index rows() const { return m_rows; }
  Index cols() const { return m_cols; }
  Index nonZeros() const { return (Index)values.size(); }
  const StorageIndex* outerIndexPtr() const { return outer.data(); }
  const StorageIndex* innerIndexPtr() const { return inner.data(); }
  const double* valuePtr() const { return values.data(); }
  double coeff(Index i, Index j) const {
    const Index o = Major ? i : j, in = Major ? j : i;
    for (StorageIndex p = outer[o]; p < outer[o + 1]; p++)
      if (inner[p] == in) return values[p];
    return 0.0;
  }
};
enum { ColMajor = 0, RowMajor = 1 };

struct PermutationMatrix {
  std::vector<StorageIndex> m_indices;
  const std::vector<StorageIndex>& indices() const { return m_indices; }
  std::vector<StorageIndex>& indices() { return m_indices; }
  Index size() const { return (Index)m_indices.size(); }
  Index rows() const { return size(); }
  void setIdentity(Index n) { m_indices.resize(n); for (Index i = 0; i < n; i++) m_indices[i] = (StorageIndex)i; }
  // (P * v)(indices(i)) = v(i), as Eigen
  VectorXd operator*(const VectorXd& v) const {
    VectorXd r(v.size());
    for (size_t i = 0; i < v.size(); i++) r[m_indices[i]] = v[i];
    return r;
  }
};

// per-block dense solver tags (template parameter _BlockQRSolver of the reference)
template <typename BlockMatrix> struct HouseholderQR { using MatrixType = BlockMatrix; static constexpr int pivoting = QRK_PIVOT_NONE; };
template <typename BlockMatrix> struct ColPivHouseholderQR { using MatrixType = BlockMatrix; static constexpr int pivoting = QRK_PIVOT_COLPIV; };
// dense right-block solver of BlockAngularSparseQR without column pivoting (reference src/QRKit/BlockedThinDenseQR.h:62;
// test/test-qrkit.cpp:53-56): P2 = identity, rank() = cols; the panel width does not change R
template <typename DenseMatrix, int SuggestedBlockCols = 2> struct BlockedThinDenseQR { using MatrixType = DenseMatrix; static constexpr int pivoting = QRK_PIVOT_NONE; };
// rank-revealing right-block solver (reference src/QRKit/BlockedThinSparseQR.h:105-283; test/test-qrkit.cpp:54-57): ColPiv inside
// panels of SuggestedBlockCols columns, zero-pivot columns deferred to the end of P2, rank() = the nonzero pivots
template <typename Matrix_, int SuggestedBlockCols = 2> struct BlockedThinSparseQR {
  using MatrixType = Matrix_;
  static constexpr int pivoting = QRK_PIVOT_COLPIV;
  static constexpr int thin_sparse_panel = SuggestedBlockCols;
};
namespace detail {
template <typename T, typename = void> struct RightSolverKind { static constexpr int kind = T::pivoting == QRK_PIVOT_COLPIV ? QRK_RIGHT_COLPIV : QRK_RIGHT_UNPIVOTED, panel = 0; };
template <typename T> struct RightSolverKind<T, decltype((void)T::thin_sparse_panel)> { static constexpr int kind = QRK_RIGHT_THIN_SPARSE, panel = T::thin_sparse_panel; };
}  // namespace detail

// ---- SparseBlockDiagonal (SparseBlockDiagonal.h:44-163) ------------------------------------------------
template <typename BlockMatrixType>
class SparseBlockDiagonal {
 public:
  using Scalar = double;
  using BlockVec = std::vector<BlockMatrixType>;
  SparseBlockDiagonal() {}
  SparseBlockDiagonal(Index rows, Index cols) : nRows(rows), nCols(cols) {}
  void insertBack(const BlockMatrixType& b) { blocks.push_back(b); }            // :141-143
  void reserve(Index n) { blocks.reserve((size_t)n); }
  Index size() const { return (Index)blocks.size(); }
  Index rows() const { return nRows; }
  Index cols() const { return nCols; }
  void setDims(Index r, Index c) { nRows = r; nCols = c; }
  const BlockMatrixType& operator[](Index i) const { return blocks[(size_t)i]; }
  BlockMatrixType& operator[](Index i) { return blocks[(size_t)i]; }
  // fromBlockDiagonalPattern (:72-89) for a column-major compressed sparse matrix with equal blocks at (i*br, i*bc)
  void fromBlockDiagonalPattern(const SparseMatrix<ColMajor>& mat, Index blockRows, Index blockCols) {
    nRows = mat.rows(); nCols = mat.cols();
    const Index numBlocks = nCols / blockCols;                                   // SparseQRUtils.h:260
    blocks.assign((size_t)numBlocks, BlockMatrixType());
    for (Index i = 0; i < numBlocks; i++)
      for (Index j = 0; j < blockCols; j++)
        for (Index k = 0; k < blockRows; k++) blocks[(size_t)i](k, j) = mat.coeff(i * blockRows + k, i * blockCols + j);
  }
  const BlockVec& blockVector() const { return blocks; }

 private:
  BlockVec blocks;
  Index nRows = 0, nCols = 0;
};

namespace detail {
// Every failing status throws, QRK_STATUS_NO_DEVICE included: there is no CPU fallback, so a compute()/solve() that cannot
// run must not return as if it had (the handle-creating helpers probe for a device on their own and record it in info()).
inline void throw_if(int status, qrk_handle_t h, const char* what) {
  if (status != QRK_STATUS_OK)
    throw std::runtime_error(std::string(what) + ": " + qrk_status_string(status) + (h ? std::string(" — ") + qrk_last_error(h) : ""));
}
inline void require(bool ok, const char* what) {
  if (!ok) throw std::invalid_argument(what);
}
}  // namespace detail

// ---- BlockDiagonalSparseQR (BlockDiagonalSparseQR.h:37-335) ---------------------------------------------
template <typename _BlockQRSolver, int _QFormat = 0>
class BlockDiagonalSparseQR {
 public:
  using BlockQRSolver = _BlockQRSolver;
  using BlockMatrixType = typename BlockQRSolver::MatrixType;
  using MatrixType = SparseBlockDiagonal<BlockMatrixType>;
  using Scalar = double;
  using MatrixQType = SparseMatrix<RowMajor>;
  using MatrixRType = SparseMatrix<ColMajor>;
  using PermutationType = PermutationMatrix;
  enum MatrixQFormat { FullQ = 0, BlockDiagonalQ = 1 };

  BlockDiagonalSparseQR() {}
  explicit BlockDiagonalSparseQR(const MatrixType& mat) { compute(mat); }
  ~BlockDiagonalSparseQR() { qrk_destroy(m_h); }
  BlockDiagonalSparseQR(const BlockDiagonalSparseQR&) = delete;
  BlockDiagonalSparseQR& operator=(const BlockDiagonalSparseQR&) = delete;

  void compute(const MatrixType& mat, const PermutationType& rowPerm = PermutationType(), bool = false) {   // :94-102
    analyzePattern(mat, rowPerm);
    m_isInitialized = false;
    factorize(mat);
  }
  void analyzePattern(const MatrixType& mat, const PermutationType& rowPerm = PermutationType()) {           // :392-405
    ensureHandle(mat);
    if (!m_h) return;
    detail::throw_if(qrk_analyze_pattern(m_h, rowPerm.size() ? rowPerm.indices().data() : nullptr), m_h, "analyzePattern");
  }
  void factorize(const MatrixType& mat) {                                                                     // :415-547
    ensureHandle(mat);
    if (!m_h) detail::throw_if(QRK_STATUS_NO_DEVICE, nullptr, "factorize");
    static_assert(sizeof(BlockMatrixType) == sizeof(double) * BlockMatrixType::RowsAtCompileTime * BlockMatrixType::ColsAtCompileTime,
                  "fixed-size blocks are stored back to back: the std::vector of blocks IS the block-COO value array");
    const double* values = mat.size() ? mat[0].data() : nullptr;
    detail::throw_if(qrk_set_blocks(m_h, values, QRK_HOST), m_h, "factorize/upload");
    detail::throw_if(qrk_factorize(m_h), m_h, "factorize");
    m_R = MatrixRType(); m_haveR = false;
    m_isInitialized = true;
  }
  Index rows() const { return m_rows; }
  Index cols() const { return m_cols; }
  const MatrixRType& matrixR() const {                                                                        // :156
    assert(m_isInitialized && "Decomposition is not initialized.");
    if (!m_haveR) {
      int64_t nnz = 0;
      detail::throw_if(qrk_matrix_r_nnz(m_h, &nnz), m_h, "matrixR");
      m_R.m_rows = m_rows; m_R.m_cols = m_cols;
      m_R.outer.resize((size_t)m_cols + 1); m_R.inner.resize((size_t)nnz); m_R.values.resize((size_t)nnz);
      detail::throw_if(qrk_matrix_r(m_h, m_R.outer.data(), m_R.inner.data(), m_R.values.data(), QRK_HOST), m_h, "matrixR");
      m_haveR = true;
    }
    return m_R;
  }
  Index rank() const {                                                                                        // :161-165
    assert(m_isInitialized && "The factorization should be called first, use compute()");
    int64_t r = 0;
    qrk_rank(m_h, &r);
    return r;
  }

  // matrixQ(): an operator with transpose() and operator*; the explicit sparse matrix on request (toSparse()).
  class QProxy {
   public:
    QProxy(const BlockDiagonalSparseQR& qr, bool transposed) : m_qr(qr), m_t(transposed) {}
    QProxy transpose() const { return QProxy(m_qr, !m_t); }
    QProxy adjoint() const { return transpose(); }
    Index rows() const { return m_qr.rows(); }
    Index cols() const { return m_qr.rows(); }
    VectorXd operator*(const VectorXd& v) const {
      VectorXd y((size_t)m_qr.rows());
      detail::throw_if((m_t ? qrk_apply_qt : qrk_apply_q)(m_qr.m_h, v.data(), m_qr.rows(), y.data(), m_qr.rows(), 1, QRK_HOST), m_qr.m_h, "matrixQ()*v");
      return y;
    }
    MatrixXd operator*(const MatrixXd& B) const {
      MatrixXd Y(m_qr.rows(), B.cols());
      detail::throw_if((m_t ? qrk_apply_qt : qrk_apply_q)(m_qr.m_h, B.data(), B.rows(), Y.data(), Y.rows(), (int32_t)B.cols(), QRK_HOST), m_qr.m_h, "matrixQ()*B");
      return Y;
    }
    MatrixQType toSparse() const {             // the reference's explicit row-major sparse Q (:455-470, 483-491, 530-533)
      MatrixQType Q;
      int64_t nnz = 0;
      detail::throw_if(qrk_matrix_q_nnz(m_qr.m_h, &nnz), m_qr.m_h, "matrixQ");
      Q.m_rows = Q.m_cols = m_qr.rows();
      Q.outer.resize((size_t)m_qr.rows() + 1); Q.inner.resize((size_t)nnz); Q.values.resize((size_t)nnz);
      detail::throw_if(qrk_matrix_q(m_qr.m_h, Q.outer.data(), Q.inner.data(), Q.values.data(), QRK_HOST), m_qr.m_h, "matrixQ");
      return Q;
    }
   private:
    const BlockDiagonalSparseQR& m_qr;
    bool m_t;
  };
  QProxy matrixQ() const { return QProxy(*this, false); }                                                     // :235-237

  const PermutationType& colsPermutation() const {                                                            // :242-246
    assert(m_isInitialized && "Decomposition is not initialized.");
    m_outputPerm_c.indices().resize((size_t)m_cols);
    detail::throw_if(qrk_cols_permutation(m_h, m_outputPerm_c.indices().data(), QRK_HOST), m_h, "colsPermutation");
    return m_outputPerm_c;
  }
  const PermutationType& rowsPermutation() const {                                                            // :251-254
    assert(m_isInitialized && "Decomposition is not initialized.");
    m_rowPerm.indices().resize((size_t)m_rows);
    detail::throw_if(qrk_rows_permutation(m_h, m_rowPerm.indices().data(), QRK_HOST), m_h, "rowsPermutation");
    return m_rowPerm;
  }
  VectorXd solve(const VectorXd& B) const {                                                                   // :258-299
    assert(m_isInitialized && "The factorization should be called first, use compute()");
    assert(rows() == (Index)B.size() && "SparseQR::solve() : invalid number of rows in the right hand side matrix");
    VectorXd x((size_t)m_cols);
    detail::throw_if(qrk_solve(m_h, B.data(), m_rows, x.data(), m_cols, 1, QRK_HOST), m_h, "solve");
    return x;
  }
  MatrixXd solve(const MatrixXd& B) const {
    MatrixXd X(m_cols, B.cols());
    detail::throw_if(qrk_solve(m_h, B.data(), B.rows(), X.data(), m_cols, (int32_t)B.cols(), QRK_HOST), m_h, "solve");
    return X;
  }
  // one fused pass: compute(mat) followed by solve(b)
  VectorXd computeAndSolve(const MatrixType& mat, const VectorXd& b) {
    ensureHandle(mat);
    VectorXd x((size_t)m_cols);
    if (!m_h) detail::throw_if(QRK_STATUS_NO_DEVICE, nullptr, "computeAndSolve");
    detail::throw_if(qrk_compute_solve(m_h, mat.size() ? mat[0].data() : nullptr, b.data(), x.data(), QRK_HOST), m_h, "computeAndSolve");
    m_haveR = false; m_isInitialized = true;
    return x;
  }
  ComputationInfo info() const {                                                                              // :309-313
    if (!m_h) return InvalidInput;
    int32_t i = 0;
    qrk_info(m_h, &i);
    return (ComputationInfo)i;
  }
  std::string lastErrorMessage() const { return m_h ? qrk_last_error(m_h) : m_lastError; }
  qrk_handle_t handle() const { return m_h; }

 protected:
  void ensureHandle(const MatrixType& mat) {
    const Index br = BlockMatrixType::RowsAtCompileTime, bc = BlockMatrixType::ColsAtCompileTime;
    if (m_h && m_nb == mat.size() && m_rows == mat.rows() && m_cols == mat.cols()) return;
    qrk_destroy(m_h);
    m_h = nullptr;
    qrk_desc_t d{};
    d.kind = QRK_BLOCK_DIAGONAL; d.num_blocks = mat.size(); d.block_rows = (int32_t)br; d.block_cols = (int32_t)bc;
    d.n_rows = mat.rows(); d.n_cols = mat.cols(); d.pivoting = BlockQRSolver::pivoting; d.q_format = _QFormat;
    const int st = qrk_create(&d, &m_h);
    m_nb = mat.size(); m_rows = mat.rows() ? mat.rows() : mat.size() * br; m_cols = mat.cols() ? mat.cols() : mat.size() * bc;
    if (st == QRK_STATUS_NO_DEVICE) { m_lastError = qrk_status_string(st); m_h = nullptr; return; }
    detail::throw_if(st, nullptr, "BlockDiagonalSparseQR");
  }
  qrk_handle_t m_h = nullptr;
  Index m_nb = -1, m_rows = 0, m_cols = 0;
  bool m_isInitialized = false;
  mutable bool m_haveR = false;
  mutable MatrixRType m_R;
  mutable PermutationType m_outputPerm_c, m_rowPerm;
  std::string m_lastError;
};

// ---- BlockMatrix1x2 (BlockMatrix1x2.h:31-67) -----------------------------------------------------------------
template <typename LeftBlock, typename RightBlock>
class BlockMatrix1x2 {
 public:
  BlockMatrix1x2(const LeftBlock& l, const RightBlock& r) : m_left(l), m_right(r) {
    assert(l.rows() == r.rows() && "blocks must have the same number of rows");                               // :37
  }
  Index rows() const { return m_left.rows(); }
  Index cols() const { return m_left.cols() + m_right.cols(); }
  const LeftBlock& leftBlock() const { return m_left; }
  const RightBlock& rightBlock() const { return m_right; }
 private:
  const LeftBlock& m_left;       // references, as the reference (:65-66): both blocks must outlive this object
  const RightBlock& m_right;
};

// ---- BlockAngularSparseQR (BlockAngularSparseQR.h:79-281), Left = BlockDiagonalSparseQR<...>, Right = dense ColPiv (default)
// or BlockedThinDenseQR / HouseholderQR (unpivoted).  Any border width: up to 8 columns with a ColPiv right solver take the fused
// TSQR kernels, everything else the dense right-block kernels.
template <typename BlockQRSolverLeftTag, typename RightSolverTag = ColPivHouseholderQR<MatrixXd>>
class BlockAngularSparseQR {
 public:
  using LeftBlockMatrixType = SparseBlockDiagonal<typename BlockQRSolverLeftTag::MatrixType>;
  using RightBlockMatrixType = MatrixXd;
  using MatrixType = BlockMatrix1x2<LeftBlockMatrixType, RightBlockMatrixType>;
  using MatrixRType = SparseMatrix<ColMajor>;
  using PermutationType = PermutationMatrix;

  BlockAngularSparseQR() {}
  explicit BlockAngularSparseQR(const MatrixType& mat) { compute(mat); }
  ~BlockAngularSparseQR() { qrk_destroy(m_h); }
  BlockAngularSparseQR(const BlockAngularSparseQR&) = delete;
  BlockAngularSparseQR& operator=(const BlockAngularSparseQR&) = delete;

  void compute(const MatrixType& mat) {                                                                       // :134-138
    ensureHandle(mat);
    if (!m_h) detail::throw_if(QRK_STATUS_NO_DEVICE, nullptr, "compute");
    const auto& L = mat.leftBlock();
    detail::throw_if(qrk_set_border(m_h, mat.rightBlock().data(), mat.rightBlock().rows(), QRK_HOST), m_h, "compute/border");
    detail::throw_if(qrk_compute(m_h, L.size() ? L[0].data() : nullptr, QRK_HOST), m_h, "compute");
    m_haveR = false; m_isInitialized = true;
  }
  VectorXd computeAndSolve(const MatrixType& mat, const VectorXd& b) {
    ensureHandle(mat);
    VectorXd x((size_t)mat.cols());
    if (!m_h) detail::throw_if(QRK_STATUS_NO_DEVICE, nullptr, "computeAndSolve");
    const auto& L = mat.leftBlock();
    detail::throw_if(qrk_set_border(m_h, mat.rightBlock().data(), mat.rightBlock().rows(), QRK_HOST), m_h, "border");
    detail::throw_if(qrk_compute_solve(m_h, L.size() ? L[0].data() : nullptr, b.data(), x.data(), QRK_HOST), m_h, "computeAndSolve");
    m_haveR = false; m_isInitialized = true;
    return x;
  }
  Index rows() const { return m_rows; }
  Index cols() const { return m_cols; }
  Index leftBlockRows() const { return m_rows; }                                                              // :276-280
  Index leftBlockCols() const { return m_cols - m_m2; }
  const MatrixRType& matrixR() const {                                                                        // :163
    if (!m_haveR) {
      int64_t nnz = 0;
      detail::throw_if(qrk_matrix_r_nnz(m_h, &nnz), m_h, "matrixR");
      m_R.m_rows = m_rows; m_R.m_cols = m_cols;
      m_R.outer.resize((size_t)m_cols + 1); m_R.inner.resize((size_t)nnz); m_R.values.resize((size_t)nnz);
      detail::throw_if(qrk_matrix_r(m_h, m_R.outer.data(), m_R.inner.data(), m_R.values.data(), QRK_HOST), m_h, "matrixR");
      m_haveR = true;
    }
    return m_R;
  }
  Index rank() const { int64_t r = 0; qrk_rank(m_h, &r); return r; }                                         // :510
  const PermutationType& colsPermutation() const {
    m_outputPerm_c.indices().resize((size_t)m_cols);
    detail::throw_if(qrk_cols_permutation(m_h, m_outputPerm_c.indices().data(), QRK_HOST), m_h, "colsPermutation");
    return m_outputPerm_c;
  }
  const PermutationType& rowsPermutation() const {
    m_rowPerm.setIdentity(m_rows);                                                                            // :440-442, identity left and right
    return m_rowPerm;
  }
  VectorXd solve(const VectorXd& B) const {                                                                   // :203-227
    assert(m_isInitialized && "The factorization should be called first, use compute()");
    VectorXd x((size_t)m_cols);
    detail::throw_if(qrk_solve(m_h, B.data(), m_rows, x.data(), m_cols, 1, QRK_HOST), m_h, "solve");
    return x;
  }
  void setPivotThreshold(double) {}                                                                           // no-op in the reference too (:234-237)
  ComputationInfo info() const { if (!m_h) return InvalidInput; int32_t i = 0; qrk_info(m_h, &i); return (ComputationInfo)i; }
  std::string lastErrorMessage() const { return m_h ? qrk_last_error(m_h) : m_lastError; }
  qrk_handle_t handle() const { return m_h; }

 private:
  void ensureHandle(const MatrixType& mat) {
    using Blk = typename BlockQRSolverLeftTag::MatrixType;
    const auto& L = mat.leftBlock();
    const Index m2 = mat.rightBlock().cols();
    assert(L.cols() > m2 && "the left block should be the bigger one");                                        // :434
    if (m_h && m_nb == L.size() && m_m2 == m2) return;
    qrk_destroy(m_h);
    m_h = nullptr;
    qrk_desc_t d{};
    d.kind = QRK_BLOCK_ANGULAR; d.num_blocks = L.size(); d.block_rows = Blk::RowsAtCompileTime; d.block_cols = Blk::ColsAtCompileTime;
    d.pivoting = BlockQRSolverLeftTag::pivoting; d.q_format = QRK_FULL_Q; d.border_cols = (int32_t)m2;
    d.right_solver = detail::RightSolverKind<RightSolverTag>::kind;
    d.reserved[1] = detail::RightSolverKind<RightSolverTag>::panel;
    const int st = qrk_create(&d, &m_h);
    m_nb = L.size(); m_m2 = m2; m_rows = mat.rows(); m_cols = mat.cols();
    if (st == QRK_STATUS_NO_DEVICE) { m_lastError = qrk_status_string(st); m_h = nullptr; return; }
    detail::throw_if(st, nullptr, "BlockAngularSparseQR");
  }
  qrk_handle_t m_h = nullptr;
  Index m_nb = -1, m_m2 = 0, m_matCols = 0, m_rows = 0, m_cols = 0;
  bool m_isInitialized = false;
  mutable bool m_haveR = false;
  mutable MatrixRType m_R;
  mutable PermutationType m_outputPerm_c, m_rowPerm;
  std::string m_lastError;
};

// ---- BandedBlockedSparseQR (BandedBlockedSparseQR.h:122-344) for the fixed block-banded pattern of fromBlockBandedPattern
// (SparseQRUtils.h:274-302): num_blocks dense BlockRows x BlockCols slabs, slab k at rows [k*BlockRows, ...), columns
// [k*(BlockCols-BlockOverlap), ...).  The input is the block-COO array of the slabs (column-major, back to back).
// Single GPU (groups of slabs are reduced in parallel, a short chase across the group boundaries stays sequential).
// matrixQ() is available in operator form on the thin part: applyQt (its transpose-apply, as the reference uses it, :299) and applyQ.
template <int BlockRows, int BlockCols, int BlockOverlap>
class BandedBlockedSparseQR {
 public:
  using MatrixRType = SparseMatrix<ColMajor>;
  using PermutationType = PermutationMatrix;
  BandedBlockedSparseQR() {}
  ~BandedBlockedSparseQR() { qrk_destroy(m_h); }
  BandedBlockedSparseQR(const BandedBlockedSparseQR&) = delete;
  BandedBlockedSparseQR& operator=(const BandedBlockedSparseQR&) = delete;

  // matCols: columns of the matrix when its last slab is narrower than BlockCols — the reference's pattern gives the last
  // block BlockCols - BlockOverlap columns (SparseQRUtils.h:284, test/test-qrkit.cpp:63-96); 0 = a full last slab
  void compute(const std::vector<double>& slabs, Index numBlocks, Index matCols = 0) {                        // :170-178
    ensureHandle(numBlocks, matCols);
    detail::require((Index)slabs.size() >= numBlocks * BlockRows * BlockCols, "BandedBlockedSparseQR::compute: slabs holds fewer than numBlocks * BlockRows * BlockCols values");
    detail::throw_if(m_h ? qrk_compute(m_h, slabs.data(), QRK_HOST) : (int)QRK_STATUS_NO_DEVICE, m_h, "compute");
    m_haveR = false; m_isInitialized = true;
  }
  // compute(const MatrixType&) on a GENERAL banded sparse matrix — the analyzePattern else-branch of the reference
  // (BandedBlockedSparseQR.h:408-426): AsBandedAsPossible row ordering, block detection, dense extraction of the blocks, then
  // the general window chain (any block sizes; BlockRows / BlockCols / BlockOverlap of this class are not used).
  // rowsPermutation() holds the ordering; as in the reference, solve() expects the right-hand side already permuted
  // (test/test-qrkit.cpp:235).  matrixQ() is an exact n x n operator on this path: applyQtFull / applyQFull.
  void compute(const SparseMatrix<ColMajor>& mat) {
    const Index n = mat.rows(), m = mat.cols();
    // row-major pattern of the matrix
    std::vector<StorageIndex> rp((size_t)n + 1, 0), ri((size_t)mat.nonZeros());
    for (Index p = 0; p < mat.nonZeros(); p++) rp[(size_t)mat.inner[(size_t)p] + 1]++;
    for (Index i = 0; i < n; i++) rp[(size_t)i + 1] += rp[(size_t)i];
    { std::vector<StorageIndex> fill(rp.begin(), rp.end() - 1);
      for (Index j = 0; j < m; j++) for (StorageIndex p = mat.outer[(size_t)j]; p < mat.outer[(size_t)j + 1]; p++) ri[(size_t)fill[(size_t)mat.inner[(size_t)p]]++] = (StorageIndex)j; }
    m_rowPerm.indices().resize((size_t)n);
    int32_t has = 0;
    detail::throw_if(qrk_order_as_banded_as_possible(n, m, rp.data(), ri.data(), m_rowPerm.indices().data(), &has), nullptr, "AsBandedAsPossible");
    // pattern of P * A (new row = indices[orig row]), then one block per band start
    std::vector<StorageIndex> prp((size_t)n + 1, 0), pri(ri.size());
    std::vector<StorageIndex> inv((size_t)n);
    for (Index i = 0; i < n; i++) inv[(size_t)m_rowPerm.indices()[(size_t)i]] = (StorageIndex)i;
    for (Index i = 0; i < n; i++) prp[(size_t)i + 1] = prp[(size_t)i] + (rp[(size_t)inv[(size_t)i] + 1] - rp[(size_t)inv[(size_t)i]]);
    for (Index i = 0; i < n; i++) std::copy(ri.begin() + rp[(size_t)inv[(size_t)i]], ri.begin() + rp[(size_t)inv[(size_t)i] + 1], pri.begin() + prp[(size_t)i]);
    int64_t nblk = 0;
    detail::throw_if(qrk_detect_band_starts(n, m, prp.data(), pri.data(), nullptr, 0, &nblk), nullptr, "block detection");
    std::vector<int32_t> blocks((size_t)nblk * 4);
    detail::throw_if(qrk_detect_band_starts(n, m, prp.data(), pri.data(), blocks.data(), nblk, &nblk), nullptr, "block detection");
    size_t total = 0;
    for (int64_t k = 0; k < nblk; k++) total += (size_t)blocks[4 * k + 2] * blocks[4 * k + 3];
    std::vector<double> values(total);
    detail::throw_if(qrk_extract_blocks(n, m, mat.outer.data(), mat.inner.data(), mat.values.data(), m_rowPerm.indices().data(), blocks.data(), nblk, values.data()),
                     nullptr, "block extraction");
    qrk_destroy(m_h);
    m_h = nullptr; m_nb = -1;
    detail::throw_if(qrk_create_banded_general(blocks.data(), nblk, n, m, 0, 0, &m_h), nullptr, "BandedBlockedSparseQR (general sparse matrix)");
    m_rows = n; m_cols = m;
    detail::throw_if(qrk_compute(m_h, values.data(), QRK_HOST), m_h, "compute");
    detail::throw_if(qrk_analyze_pattern(m_h, m_rowPerm.indices().data()), m_h, "rowsPermutation");
    m_haveR = false; m_isInitialized = true;
  }
  const PermutationType& rowsPermutation() const { return m_rowPerm; }
  VectorXd applyQtFull(const VectorXd& v) const {                                                             // matrixQ().transpose() * v, n x n (general path)
    detail::require((Index)v.size() == m_rows, "applyQtFull: v.size() != rows()");
    VectorXd y((size_t)m_rows);
    detail::throw_if(qrk_apply_qt(m_h, v.data(), m_rows, y.data(), m_rows, 1, QRK_HOST), m_h, "matrixQ().transpose() * v");
    return y;
  }
  VectorXd applyQFull(const VectorXd& v) const {                                                              // matrixQ() * v, n x n (general path)
    detail::require((Index)v.size() == m_rows, "applyQFull: v.size() != rows()");
    VectorXd y((size_t)m_rows);
    detail::throw_if(qrk_apply_q(m_h, v.data(), m_rows, y.data(), m_rows, 1, QRK_HOST), m_h, "matrixQ() * v");
    return y;
  }
  VectorXd computeAndSolve(const std::vector<double>& slabs, Index numBlocks, const VectorXd& b, Index matCols = 0) {
    ensureHandle(numBlocks, matCols);
    detail::require((Index)slabs.size() >= numBlocks * BlockRows * BlockCols, "BandedBlockedSparseQR::computeAndSolve: slabs holds fewer than numBlocks * BlockRows * BlockCols values");
    detail::require((Index)b.size() == m_rows, "BandedBlockedSparseQR::computeAndSolve: b.size() != rows()");
    VectorXd x((size_t)m_cols);
    detail::throw_if(m_h ? qrk_compute_solve(m_h, slabs.data(), b.data(), x.data(), QRK_HOST) : (int)QRK_STATUS_NO_DEVICE, m_h, "computeAndSolve");
    m_haveR = false; m_isInitialized = true;
    return x;
  }
  Index rows() const { return m_rows; }
  Index cols() const { return m_cols; }
  Index rank() const { int64_t r = 0; qrk_rank(m_h, &r); return r; }                                         // :514
  ComputationInfo info() const { if (!m_h) return InvalidInput; int32_t i = 0; qrk_info(m_h, &i); return (ComputationInfo)i; }
  std::string lastErrorMessage() const { return m_h ? qrk_last_error(m_h) : m_lastError; }
  const MatrixRType& matrixR() const {                                                                        // :484-491, explicit zeros included
    if (!m_haveR) {
      int64_t nnz = 0;
      detail::throw_if(qrk_matrix_r_nnz(m_h, &nnz), m_h, "matrixR");
      m_R.m_rows = m_rows; m_R.m_cols = m_cols;
      m_R.outer.resize((size_t)m_cols + 1); m_R.inner.resize((size_t)nnz); m_R.values.resize((size_t)nnz);
      detail::throw_if(qrk_matrix_r(m_h, m_R.outer.data(), m_R.inner.data(), m_R.values.data(), QRK_HOST), m_h, "matrixR");
      m_haveR = true;
    }
    return m_R;
  }
  // The thin factor Q1 = A R^-1 (rows() x cols()): what solve() and the LM caller read of matrixQ() (:299, topRows(rank)).
  // The two-phase factorisation has no n x n Q (banded.cuh), so these are the Q products of this solver.
  VectorXd applyQt(const VectorXd& v) const {                                                                 // (matrixQ().transpose() * v).topRows(cols()) (:655-670)
    detail::require((Index)v.size() == m_rows, "applyQt: v.size() != rows()");
    VectorXd y((size_t)m_cols);
    detail::throw_if(qrk_apply_qt_thin(m_h, v.data(), m_rows, y.data(), m_cols, 1, QRK_HOST), m_h, "matrixQ().transpose() * v");
    return y;
  }
  VectorXd applyQ(const VectorXd& y) const {                                                                  // matrixQ() * [y; 0] (:640-675)
    detail::require((Index)y.size() == m_cols, "applyQ: y.size() != cols()");
    VectorXd x((size_t)m_rows);
    detail::throw_if(qrk_apply_q_thin(m_h, y.data(), m_cols, x.data(), m_rows, 1, QRK_HOST), m_h, "matrixQ() * y");
    return x;
  }
  VectorXd solve(const VectorXd& B) const {                                                                   // :287-307
    assert(m_isInitialized && "The factorization should be called first, use compute()");
    detail::require((Index)B.size() == m_rows, "solve: B.size() != rows()");
    VectorXd x((size_t)m_cols);
    detail::throw_if(qrk_solve(m_h, B.data(), m_rows, x.data(), m_cols, 1, QRK_HOST), m_h, "solve");
    return x;
  }
 private:
  void ensureHandle(Index numBlocks, Index matCols) {
    if (m_h && m_nb == numBlocks && m_matCols == matCols) return;
    qrk_destroy(m_h);
    m_h = nullptr;
    qrk_desc_t d{};
    d.kind = QRK_BANDED_BLOCKED; d.num_blocks = numBlocks; d.block_rows = BlockRows; d.block_cols = BlockCols; d.block_overlap = BlockOverlap;
    d.n_cols = matCols;
    const int st = qrk_create(&d, &m_h);
    m_nb = numBlocks; m_matCols = matCols;
    if (st == QRK_STATUS_NO_DEVICE) { m_lastError = qrk_status_string(st); m_h = nullptr; return; }
    detail::throw_if(st, nullptr, "BandedBlockedSparseQR");
    int64_t r = 0, c = 0;
    qrk_rows(m_h, &r); qrk_cols(m_h, &c);
    m_rows = r; m_cols = c;
  }
  qrk_handle_t m_h = nullptr;
  Index m_nb = -1, m_matCols = 0, m_rows = 0, m_cols = 0;
  bool m_isInitialized = false;
  mutable bool m_haveR = false;
  mutable MatrixRType m_R;
  PermutationType m_rowPerm;
  std::string m_lastError;
};

// ---- BlockAngularSparseQR<BandedBlockedSparseQR<..., BlockOverlap, ...>, RightSolver> (the solver pair of the reference's own
// block-angular tests, test/test-qrkit.cpp:44-57, and of the QRkitBB benchmark column): the left block J1 is block banded and
// given as its slabs (see BandedBlockedSparseQR above), the right block is dense (n x m2, column-major).
template <int BlockRows, int BlockCols, int BlockOverlap, typename RightSolverTag = ColPivHouseholderQR<MatrixXd>>
class BlockAngularBandedSparseQR {
 public:
  using MatrixRType = SparseMatrix<ColMajor>;
  using PermutationType = PermutationMatrix;
  BlockAngularBandedSparseQR() {}
  ~BlockAngularBandedSparseQR() { qrk_destroy(m_h); }
  BlockAngularBandedSparseQR(const BlockAngularBandedSparseQR&) = delete;
  BlockAngularBandedSparseQR& operator=(const BlockAngularBandedSparseQR&) = delete;

  // matCols: columns of the whole matrix [J1 | J2] when the last slab of J1 is narrower than BlockCols (see BandedBlockedSparseQR)
  void compute(const std::vector<double>& slabs, Index numBlocks, const MatrixXd& border, Index matCols = 0) {   // :134-138
    ensureHandle(numBlocks, border, matCols);
    detail::require((Index)slabs.size() >= numBlocks * BlockRows * BlockCols, "BlockAngularBandedSparseQR::compute: slabs holds fewer than numBlocks * BlockRows * BlockCols values");
    detail::require(border.rows() == m_rows, "BlockAngularBandedSparseQR::compute: border.rows() != rows()");
    if (!m_h) detail::throw_if(QRK_STATUS_NO_DEVICE, nullptr, "compute");
    detail::throw_if(qrk_set_border(m_h, border.data(), border.rows(), QRK_HOST), m_h, "compute/border");
    detail::throw_if(qrk_compute(m_h, slabs.data(), QRK_HOST), m_h, "compute");
    m_haveR = false; m_isInitialized = true;
  }
  VectorXd computeAndSolve(const std::vector<double>& slabs, Index numBlocks, const MatrixXd& border, const VectorXd& b, Index matCols = 0) {
    ensureHandle(numBlocks, border, matCols);
    detail::require((Index)slabs.size() >= numBlocks * BlockRows * BlockCols, "BlockAngularBandedSparseQR::computeAndSolve: slabs holds fewer than numBlocks * BlockRows * BlockCols values");
    detail::require(border.rows() == m_rows && (Index)b.size() == m_rows, "BlockAngularBandedSparseQR::computeAndSolve: border.rows() / b.size() != rows()");
    VectorXd x((size_t)m_cols);
    if (!m_h) detail::throw_if(QRK_STATUS_NO_DEVICE, nullptr, "computeAndSolve");
    detail::throw_if(qrk_set_border(m_h, border.data(), border.rows(), QRK_HOST), m_h, "border");
    detail::throw_if(qrk_compute_solve(m_h, slabs.data(), b.data(), x.data(), QRK_HOST), m_h, "computeAndSolve");
    m_haveR = false; m_isInitialized = true;
    return x;
  }
  Index rows() const { return m_rows; }
  Index cols() const { return m_cols; }
  Index leftBlockCols() const { return m_cols - m_m2; }                                                       // :276-280
  Index rank() const { int64_t r = 0; qrk_rank(m_h, &r); return r; }                                         // :510
  ComputationInfo info() const { if (!m_h) return InvalidInput; int32_t i = 0; qrk_info(m_h, &i); return (ComputationInfo)i; }
  std::string lastErrorMessage() const { return m_h ? qrk_last_error(m_h) : m_lastError; }
  const PermutationType& colsPermutation() const {                                                            // [identity ; m1 + P2] (:498-503)
    m_outputPerm_c.indices().resize((size_t)m_cols);
    detail::throw_if(qrk_cols_permutation(m_h, m_outputPerm_c.indices().data(), QRK_HOST), m_h, "colsPermutation");
    return m_outputPerm_c;
  }
  const MatrixRType& matrixR() const {                                                                        // [R1 band, Atop P2; 0, R2] (:285-308)
    if (!m_haveR) {
      int64_t nnz = 0;
      detail::throw_if(qrk_matrix_r_nnz(m_h, &nnz), m_h, "matrixR");
      m_R.m_rows = m_rows; m_R.m_cols = m_cols;
      m_R.outer.resize((size_t)m_cols + 1); m_R.inner.resize((size_t)nnz); m_R.values.resize((size_t)nnz);
      detail::throw_if(qrk_matrix_r(m_h, m_R.outer.data(), m_R.inner.data(), m_R.values.data(), QRK_HOST), m_h, "matrixR");
      m_haveR = true;
    }
    return m_R;
  }
  VectorXd solve(const VectorXd& B) const {                                                                   // :203-227
    assert(m_isInitialized && "The factorization should be called first, use compute()");
    VectorXd x((size_t)m_cols);
    detail::throw_if(qrk_solve(m_h, B.data(), m_rows, x.data(), m_cols, 1, QRK_HOST), m_h, "solve");
    return x;
  }
 private:
  void ensureHandle(Index numBlocks, const MatrixXd& border, Index matCols) {
    const Index m2 = border.cols();
    if (m_h && m_nb == numBlocks && m_m2 == m2 && m_matCols == matCols) return;
    qrk_destroy(m_h);
    m_h = nullptr;
    qrk_desc_t d{};
    d.kind = QRK_BLOCK_ANGULAR; d.left_solver = QRK_LEFT_BANDED_BLOCKED; d.num_blocks = numBlocks;
    d.block_rows = BlockRows; d.block_cols = BlockCols; d.block_overlap = BlockOverlap;
    d.pivoting = QRK_PIVOT_NONE; d.q_format = QRK_FULL_Q; d.border_cols = (int32_t)m2;
    d.right_solver = RightSolverTag::pivoting == QRK_PIVOT_COLPIV ? QRK_RIGHT_COLPIV : QRK_RIGHT_UNPIVOTED;
    d.n_cols = matCols;                                                                                     // m1 + m2 (0: full last slab)
    const int st = qrk_create(&d, &m_h);
    m_nb = numBlocks; m_m2 = m2; m_matCols = matCols; m_rows = numBlocks * BlockRows;
    m_cols = matCols > 0 ? matCols : (numBlocks - 1) * (BlockCols - BlockOverlap) + BlockCols + m2;
    if (st == QRK_STATUS_NO_DEVICE) { m_lastError = qrk_status_string(st); m_h = nullptr; return; }
    detail::throw_if(st, nullptr, "BlockAngularBandedSparseQR");
  }
  qrk_handle_t m_h = nullptr;
  Index m_nb = -1, m_m2 = 0, m_matCols = 0, m_rows = 0, m_cols = 0;
  bool m_isInitialized = false;
  mutable bool m_haveR = false;
  mutable MatrixRType m_R;
  mutable PermutationType m_outputPerm_c;
  std::string m_lastError;
};

// ---- multi-GPU, one host process -------------------------------------------------------------------------------
// The reference's block loop is serial (BlockDiagonalSparseQR.h:432); the blocks are independent, so G GPUs each take a
// contiguous range of them.  One handle per device, one host thread per handle for the (blocking) host-memory calls.
namespace detail {
// contiguous block range of shard g: balanced, every boundary a multiple of 2 blocks (16-byte alignment of odd-sized slices)
inline void shard_range(Index nb, int G, int g, Index& lo, Index& hi) {
  const Index units = (nb + 1) / 2;
  lo = std::min<Index>(nb, (units * g / G) * 2);
  hi = std::min<Index>(nb, (units * (g + 1) / G) * 2);
}
template <typename F>
inline void for_each_shard(int G, F f) {        // f(g) on its own thread; the first exception is rethrown on the caller's thread
  std::vector<std::thread> th;
  std::vector<std::exception_ptr> err((size_t)G);
  for (int g = 0; g < G; g++) th.emplace_back([&, g]() { try { f(g); } catch (...) { err[(size_t)g] = std::current_exception(); } });
  for (auto& t : th) t.join();
  for (auto& e : err) if (e) std::rethrow_exception(e);
}
}  // namespace detail

// BlockDiagonalSparseQR over several GPUs: NO collective on the data path — x, R and the permutations stay sharded on the
// devices and are concatenated on the host (global offsets = prefix sums of the block sizes).
template <typename BlockQRSolverTag>
class ShardedBlockDiagonalSparseQR {
 public:
  using Blk = typename BlockQRSolverTag::MatrixType;
  using MatrixType = SparseBlockDiagonal<Blk>;
  explicit ShardedBlockDiagonalSparseQR(const std::vector<int>& devices) : m_dev(devices), m_h(devices.size(), nullptr) {
    detail::require(!devices.empty(), "ShardedBlockDiagonalSparseQR: no devices");
  }
  ~ShardedBlockDiagonalSparseQR() { for (auto h : m_h) qrk_destroy(h); }
  ShardedBlockDiagonalSparseQR(const ShardedBlockDiagonalSparseQR&) = delete;
  ShardedBlockDiagonalSparseQR& operator=(const ShardedBlockDiagonalSparseQR&) = delete;

  void compute(const MatrixType& mat) {
    ensureHandles(mat);
    const double* values = mat.size() ? mat[0].data() : nullptr;
    detail::for_each_shard(G(), [&](int g) { detail::throw_if(qrk_compute(m_h[(size_t)g], values + m_lo[(size_t)g] * R * C, QRK_HOST), m_h[(size_t)g], "compute"); });
  }
  VectorXd solve(const VectorXd& b) const {
    detail::require((Index)b.size() == m_rows, "solve: b.size() != rows()");
    VectorXd x((size_t)m_cols);
    detail::for_each_shard(G(), [&](int g) {
      const Index lo = m_lo[(size_t)g], nbg = m_lo[(size_t)g + 1] - lo;
      detail::throw_if(qrk_solve(m_h[(size_t)g], b.data() + lo * R, nbg * R, x.data() + lo * C, nbg * C, 1, QRK_HOST), m_h[(size_t)g], "solve");
    });
    return x;
  }
  VectorXd computeAndSolve(const MatrixType& mat, const VectorXd& b) {      // the fused pass on every shard
    ensureHandles(mat);
    detail::require((Index)b.size() == m_rows, "computeAndSolve: b.size() != rows()");
    VectorXd x((size_t)m_cols);
    const double* values = mat.size() ? mat[0].data() : nullptr;
    detail::for_each_shard(G(), [&](int g) {
      const Index lo = m_lo[(size_t)g];
      detail::throw_if(qrk_compute_solve(m_h[(size_t)g], values + lo * R * C, b.data() + lo * R, x.data() + lo * C, QRK_HOST), m_h[(size_t)g], "computeAndSolve");
    });
    return x;
  }
  Index rows() const { return m_rows; }
  Index cols() const { return m_cols; }
  Index rank() const { Index r = 0; for (auto h : m_h) { int64_t rg = 0; qrk_rank(h, &rg); r += rg; } return r; }
  PermutationMatrix colsPermutation() const {                              // shard-local indices shifted by the shard's first column
    PermutationMatrix p;
    p.indices().resize((size_t)m_cols);
    for (int g = 0; g < G(); g++) {
      const Index lo = m_lo[(size_t)g], nbg = m_lo[(size_t)g + 1] - lo;
      if (nbg == 0) continue;
      detail::throw_if(qrk_cols_permutation(m_h[(size_t)g], p.indices().data() + lo * C, QRK_HOST), m_h[(size_t)g], "colsPermutation");
      for (Index j = lo * C; j < (lo + nbg) * C; j++) p.indices()[(size_t)j] += (int)(lo * C);
    }
    return p;
  }
  int shards() const { return G(); }

 private:
  enum { R = Blk::RowsAtCompileTime, C = Blk::ColsAtCompileTime };
  int G() const { return (int)m_dev.size(); }
  void ensureHandles(const MatrixType& mat) {
    if (m_nb == mat.size()) return;
    for (auto& h : m_h) { qrk_destroy(h); h = nullptr; }
    m_nb = mat.size(); m_rows = m_nb * R; m_cols = m_nb * C;
    m_lo.assign((size_t)G() + 1, 0);
    for (int g = 0; g < G(); g++) {
      Index lo, hi;
      detail::shard_range(m_nb, G(), g, lo, hi);
      m_lo[(size_t)g] = lo; m_lo[(size_t)g + 1] = hi;
      qrk_desc_t d{};
      d.kind = QRK_BLOCK_DIAGONAL; d.device = m_dev[(size_t)g]; d.num_blocks = hi - lo; d.block_rows = R; d.block_cols = C;
      d.pivoting = BlockQRSolverTag::pivoting; d.q_format = QRK_FULL_Q;
      detail::throw_if(qrk_create(&d, &m_h[(size_t)g]), nullptr, "ShardedBlockDiagonalSparseQR");
    }
  }
  std::vector<int> m_dev;
  std::vector<qrk_handle_t> m_h;
  std::vector<Index> m_lo;
  Index m_nb = -1, m_rows = 0, m_cols = 0;
};

// BlockAngularSparseQR over several GPUs (narrow border, fused TSQR path): every GPU factors its block range, applies Q1^T to
// its rows of [J2 | b] and reduces them to one m2 x (m2+1) triangle; the triangles cross NVLink INSIDE the TSQR root kernel
// (peer stores + flags, qrk_angular_p2p_attach), every GPU merges them in rank order — bit-identical shared parameters — and
// back-substitutes its own x1.  The ONE exchange step of BlockAngularSparseQR.h:361-369, no host round trip.
template <typename BlockQRSolverLeftTag>
class ShardedBlockAngularSparseQR {
 public:
  using Blk = typename BlockQRSolverLeftTag::MatrixType;
  using MatrixType = BlockMatrix1x2<SparseBlockDiagonal<Blk>, MatrixXd>;
  explicit ShardedBlockAngularSparseQR(const std::vector<int>& devices) : m_dev(devices), m_h(devices.size(), nullptr) {
    detail::require(devices.size() >= 2, "ShardedBlockAngularSparseQR: needs at least two shards (use BlockAngularSparseQR for one GPU)");
  }
  ~ShardedBlockAngularSparseQR() { for (auto h : m_h) qrk_destroy(h); }
  ShardedBlockAngularSparseQR(const ShardedBlockAngularSparseQR&) = delete;
  ShardedBlockAngularSparseQR& operator=(const ShardedBlockAngularSparseQR&) = delete;

  // x = [x1 (all shards, in block order) ; x2]
  VectorXd computeAndSolve(const MatrixType& mat, const VectorXd& b) {
    ensureHandles(mat);
    const auto& L = mat.leftBlock();
    const MatrixXd& J2 = mat.rightBlock();
    detail::require((Index)b.size() == m_rows && J2.rows() == m_rows, "computeAndSolve: b.size() / border rows != rows()");
    const double* values = L.size() ? L[0].data() : nullptr;
    VectorXd x((size_t)m_cols);
    std::vector<VectorXd> xg((size_t)G());
    detail::for_each_shard(G(), [&](int g) {
      const Index lo = m_lo[(size_t)g], nbg = m_lo[(size_t)g + 1] - lo;
      qrk_handle_t h = m_h[(size_t)g];
      xg[(size_t)g] = VectorXd((size_t)(nbg * C + m_m2));
      detail::throw_if(qrk_set_border(h, J2.data() + lo * R, J2.rows(), QRK_HOST), h, "border");      // this shard's rows of every border column
      detail::throw_if(qrk_compute_solve(h, values + lo * R * C, b.data() + lo * R, xg[(size_t)g].data(), QRK_HOST), h, "computeAndSolve");
    });
    for (int g = 0; g < G(); g++) {
      const Index lo = m_lo[(size_t)g], nbg = m_lo[(size_t)g + 1] - lo;
      std::copy(xg[(size_t)g].begin(), xg[(size_t)g].begin() + nbg * C, x.begin() + lo * C);
    }
    const Index m1 = m_cols - m_m2, last = m_lo[1] - m_lo[0];
    std::copy(xg[0].begin() + last * C, xg[0].end(), x.begin() + m1);                                  // x2: the same bits on every shard
    m_x2_identical = true;
    for (int g = 1; g < G(); g++) {
      const Index nbg = m_lo[(size_t)g + 1] - m_lo[(size_t)g];
      for (Index j = 0; j < m_m2; j++) m_x2_identical = m_x2_identical && xg[(size_t)g][(size_t)(nbg * C + j)] == x[(size_t)(m1 + j)];
    }
    return x;
  }
  bool sharedParametersIdentical() const { return m_x2_identical; }
  Index rows() const { return m_rows; }
  Index cols() const { return m_cols; }
  Index rank() const { int64_t r2 = 0, r = 0; for (int g = 0; g < G(); g++) { qrk_rank(m_h[(size_t)g], &r); r2 = r - (m_lo[(size_t)g + 1] - m_lo[(size_t)g]) * C; } return (m_cols - m_m2) + r2; }
  int shards() const { return G(); }

 private:
  enum { R = Blk::RowsAtCompileTime, C = Blk::ColsAtCompileTime };
  int G() const { return (int)m_dev.size(); }
  void ensureHandles(const MatrixType& mat) {
    const auto& L = mat.leftBlock();
    const Index m2 = mat.rightBlock().cols();
    if (m_nb == L.size() && m_m2 == m2) return;
    for (auto& h : m_h) { qrk_destroy(h); h = nullptr; }
    m_nb = L.size(); m_m2 = m2; m_rows = mat.rows(); m_cols = mat.cols();
    m_lo.assign((size_t)G() + 1, 0);
    std::vector<void*> bufs((size_t)G(), nullptr);
    for (int g = 0; g < G(); g++) {
      Index lo, hi;
      detail::shard_range(m_nb, G(), g, lo, hi);
      m_lo[(size_t)g] = lo; m_lo[(size_t)g + 1] = hi;
      qrk_desc_t d{};
      d.kind = QRK_BLOCK_ANGULAR; d.device = m_dev[(size_t)g]; d.num_blocks = hi - lo; d.block_rows = R; d.block_cols = C;
      d.pivoting = BlockQRSolverLeftTag::pivoting; d.q_format = QRK_FULL_Q; d.border_cols = (int32_t)m2; d.right_solver = QRK_RIGHT_COLPIV;
      detail::throw_if(qrk_create(&d, &m_h[(size_t)g]), nullptr, "ShardedBlockAngularSparseQR");
      detail::throw_if(qrk_angular_set_world(m_h[(size_t)g], G()), m_h[(size_t)g], "set_world (the fused TSQR path takes 1..8 border columns)");
      int64_t bytes = 0;
      detail::throw_if(qrk_angular_xchg_buffer(m_h[(size_t)g], &bufs[(size_t)g], &bytes), m_h[(size_t)g], "xchg_buffer");
    }
    for (int g = 0; g < G(); g++)
      for (int p = 0; p < G(); p++) detail::throw_if(qrk_enable_peer_access(m_dev[(size_t)g], m_dev[(size_t)p]), nullptr, "peer access between the shards' devices");
    for (int g = 0; g < G(); g++) detail::throw_if(qrk_angular_p2p_attach(m_h[(size_t)g], bufs.data(), G(), g), m_h[(size_t)g], "p2p_attach");
  }
  std::vector<int> m_dev;
  std::vector<qrk_handle_t> m_h;
  std::vector<Index> m_lo;
  Index m_nb = -1, m_m2 = 0, m_rows = 0, m_cols = 0;
  bool m_x2_identical = false;
};

}  // namespace QRKit_b200
