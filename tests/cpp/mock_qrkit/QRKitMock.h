// Mock of the three reference headers include/qrkit_b200/EigenAdapter.hpp builds on (SparseBlockDiagonal.h:44-163,
// BlockMatrix1x2.h:31-67, SparseQRUtils.h:22-35): TEST INFRASTRUCTURE ONLY, member names as in the reference.
#ifndef MOCK_QRKIT_H_
#define MOCK_QRKIT_H_
#include <Eigen/Sparse>
#include <vector>
namespace QRKit {
template <typename BlockMatrixType> struct SparseBlockDiagonal {
  typedef Eigen::Index Index;
  SparseBlockDiagonal(Index r = 0, Index c = 0) : nRows(r), nCols(c) {}
  Index rows() const { return nRows; }
  Index cols() const { return nCols; }
  Index size() const { return (Index)blocks.size(); }
  void insertBack(const BlockMatrixType& b) { blocks.push_back(b); }
  const BlockMatrixType& operator[](Index i) const { return blocks[(size_t)i]; }
  std::vector<BlockMatrixType> blocks;
  Index nRows, nCols;
};
template <typename L, typename R> struct BlockMatrix1x2 {
  typedef Eigen::Index Index;
  BlockMatrix1x2(const L& l, const R& r) : m_l(l), m_r(r) {}
  Index rows() const { return m_l.rows(); }
  Index cols() const { return m_l.cols() + m_r.cols(); }
  const L& leftBlock() const { return m_l; }
  const R& rightBlock() const { return m_r; }
  const L& m_l; const R& m_r;
};
namespace SparseQRUtils {
template <typename SolverType> struct HasRowsPermutation { static const bool value = false; };
}
}  // namespace QRKit
#endif
