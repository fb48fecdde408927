// structure.cpp — host-side pattern analysis in front of the hot path: row ordering and block-structure detection of a
// general sparse matrix, so that a caller can hand over an Eigen-style compressed matrix instead of explicit block
// descriptors (SparseBlockDiagonal::fromSparseMatrix, reference src/QRKit/SparseBlockDiagonal.h:96-130; the generic
// analyzePattern branch of BandedBlockedSparseQR, BandedBlockedSparseQR.h:417-426).
//
// Restated from the reference's behaviour (integer work, bit-exact parity is the bar; the reference's own unit tests
// hold the expected block lists, test/test-utils.cpp:182-274):
//   qrk_order_as_banded_as_possible  SparseQROrdering::AsBandedAsPossible   (SparseQROrdering.h:53-120)
//   qrk_order_column_density         SparseQROrdering::ColumnDensity        (SparseQROrdering.h:22-50)
//   qrk_detect_blocks                BlockBandedMatrixInfo::operator()      (SparseQRUtils.h:186-253) + mergeBlocks (:308-385)
//   qrk_block_diagonal_pattern       BlockBandedMatrixInfo::fromBlockDiagonalPattern (:255-272)
//   qrk_block_banded_pattern         BlockBandedMatrixInfo::fromBlockBandedPattern   (:274-302)
//   qrk_extract_blocks               the mat.block(idxRow, idxCol, numRows, numCols) loop of SparseBlockDiagonal.h:123-128
// No device code: these run on the host, as in the reference, and do not need a GPU.
#include <algorithm>
#include <cstdint>
#include <map>
#include <numeric>
#include <vector>

#include "../../include/qrkit_b200.h"

namespace {

struct Block { int32_t row, col, nrows, ncols; };

// first / last stored column of every row of a row-major pattern; an empty row starts (and ends) at `cols`
void row_extents(int64_t rows, int64_t cols, const int32_t* outer, const int32_t* inner, std::vector<int32_t>& first,
                 std::vector<int32_t>& last) {
  first.resize(rows); last.resize(rows);
  for (int64_t j = 0; j < rows; j++) {
    const int32_t b = outer[j], e = outer[j + 1];
    first[j] = (e > b) ? inner[b] : (int32_t)cols;
    last[j] = (e > b) ? inner[e - 1] : first[j];
  }
}

// Coalesce detected blocks until each is portrait, at least maxColStep and at least `suggested` columns wide;
// a block whose column range lies inside the previous output block only adds its rows to it; a remainder that cannot
// form a block of its own is absorbed by the last output block.
void merge_blocks(std::vector<Block>& blocks, int maxColStep, int suggested) {
  std::vector<Block> out;
  Block first{0, 0, 0, 0};
  int curRows = 0, curCols = 0;
  auto good = [&]() { return curRows > curCols && curCols >= maxColStep && curCols >= suggested; };
  for (const Block& b : blocks) {
    if (!out.empty()) {
      Block& lastb = out.back();
      if (b.col + b.ncols <= lastb.col + lastb.ncols) { lastb.nrows += b.nrows; continue; }
    }
    if (first.nrows == 0) { first = b; curRows = b.nrows; curCols = b.ncols; }
    else { curRows = b.row + b.nrows - first.row; curCols = b.col + b.ncols - first.col; }
    if (good()) {
      out.push_back({first.row, first.col, curRows, curCols});
      first = Block{0, 0, 0, 0};
    }
  }
  if (first.nrows != 0) {
    if (good() || out.empty()) out.push_back({first.row, first.col, curRows, curCols});
    else {
      Block& lastb = out.back();
      lastb.ncols = first.col + curCols - lastb.col;
      lastb.nrows += curRows;
    }
  }
  blocks.swap(out);
}

int emit(const std::vector<Block>& blocks, int32_t* out, int64_t capacity, int64_t* num_blocks) {
  if (num_blocks) *num_blocks = (int64_t)blocks.size();
  if (!out) return QRK_STATUS_OK;                        // size query
  if ((int64_t)blocks.size() > capacity) return QRK_STATUS_INVALID_ARGUMENT;
  for (size_t i = 0; i < blocks.size(); i++) {
    out[4 * i] = blocks[i].row; out[4 * i + 1] = blocks[i].col; out[4 * i + 2] = blocks[i].nrows; out[4 * i + 3] = blocks[i].ncols;
  }
  return QRK_STATUS_OK;
}

}  // namespace

namespace qrk {
// The windows BandedBlockedSparseQR::factorize walks (BandedBlockedSparseQR.h:463-508) for nb dense slabs of br x bc shifted by
// bc - ov columns, the last one last_cols wide: what BlockBandedMatrixInfo::operator() detects on such a matrix (one block per
// slab, SparseQRUtils.h:186-253) after mergeBlocks (:308-385) with maxColStep = bc - ov.  out4: {idxRow, idxCol, numRows, numCols}.
void banded_reference_windows(long long nb, int br, int bc, int ov, int last_cols, int suggested, std::vector<int32_t>& out4) {
  const int step = bc - ov;
  std::vector<Block> v;
  v.reserve((size_t)nb);
  for (long long k = 0; k < nb; k++)
    v.push_back({(int32_t)(k * br), (int32_t)(k * step), br, (k < nb - 1) ? bc : last_cols});
  merge_blocks(v, step, suggested);
  out4.resize(4 * v.size());
  for (size_t i = 0; i < v.size(); i++) {
    out4[4 * i] = v[i].row; out4[4 * i + 1] = v[i].col; out4[4 * i + 2] = v[i].nrows; out4[4 * i + 3] = v[i].ncols;
  }
}
// The same for a general list of detected blocks (one per distinct band start): mergeBlocks with maxColStep = the largest
// step between consecutive first columns, as BlockBandedMatrixInfo::operator() computes it (SparseQRUtils.h:213-220).
void banded_reference_windows_from_blocks(const std::vector<int32_t>& in4, int suggested, std::vector<int32_t>& out4) {
  std::vector<Block> v(in4.size() / 4);
  int maxColStep = 0;
  for (size_t i = 0; i < v.size(); i++) {
    v[i] = {in4[4 * i], in4[4 * i + 1], in4[4 * i + 2], in4[4 * i + 3]};
    if (i > 0) maxColStep = std::max(maxColStep, (int)(v[i].col - v[i - 1].col));
  }
  merge_blocks(v, maxColStep, suggested);
  out4.resize(4 * v.size());
  for (size_t i = 0; i < v.size(); i++) {
    out4[4 * i] = v[i].row; out4[4 * i + 1] = v[i].col; out4[4 * i + 2] = v[i].nrows; out4[4 * i + 3] = v[i].ncols;
  }
}
}  // namespace qrk

extern "C" {

int qrk_order_as_banded_as_possible(int64_t rows, int64_t cols, const int32_t* csr_outer, const int32_t* csr_inner,
                                    int32_t* perm_indices, int32_t* has_permutation) {
  if (rows < 0 || cols < 0 || !csr_outer || !perm_indices || (rows > 0 && csr_outer[rows] > 0 && !csr_inner)) return QRK_STATUS_INVALID_ARGUMENT;
  std::vector<int32_t> first, last;
  row_extents(rows, cols, csr_outer, csr_inner, first, last);
  std::vector<int32_t> order(rows);
  std::iota(order.begin(), order.end(), 0);
  const bool sorted = std::is_sorted(first.begin(), first.end());
  if (!sorted) std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return first[a] < first[b]; });
  for (int64_t pos = 0; pos < rows; pos++) perm_indices[order[pos]] = (int32_t)pos;   // P.indices()(orig) = new row
  if (has_permutation) *has_permutation = sorted ? 0 : 1;
  return QRK_STATUS_OK;
}

int qrk_order_column_density(int64_t cols, const int32_t* csc_outer, int32_t* perm_indices) {
  if (cols < 0 || !csc_outer || !perm_indices) return QRK_STATUS_INVALID_ARGUMENT;
  std::vector<int32_t> order(cols);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
    return (csc_outer[a + 1] - csc_outer[a]) < (csc_outer[b + 1] - csc_outer[b]);
  });
  for (int64_t pos = 0; pos < cols; pos++) perm_indices[order[pos]] = (int32_t)pos;
  return QRK_STATUS_OK;
}

static int detect_impl(int64_t rows, int64_t cols, const int32_t* csr_outer, const int32_t* csr_inner, int32_t suggested_block_cols, bool merge,
                       int32_t* blocks, int64_t capacity, int64_t* num_blocks, int64_t* nonzero_q_estimate) {
  if (rows < 0 || cols < 0 || !csr_outer) return QRK_STATUS_INVALID_ARGUMENT;
  std::vector<int32_t> first, last;
  row_extents(rows, cols, csr_outer, csr_inner, first, last);
  // widest band and number of rows per band start
  std::map<int32_t, int32_t> width, height;
  for (int64_t j = 0; j < rows; j++) {
    const int32_t bw = last[j] - first[j] + 1;
    auto it = width.find(first[j]);
    if (it == width.end()) width.emplace(first[j], bw);
    else if (it->second < bw) it->second = bw;
    height[first[j]] += 1;
  }
  int maxColStep = 0;
  for (int64_t j = 0; j + 1 < rows; j++) maxColStep = std::max(maxColStep, (int)(first[j + 1] - first[j]));
  // one block per distinct band start, in row order (the reference looks the start up with a binary search in the
  // list built so far, i.e. it relies on the starts arriving in ascending order)
  std::vector<Block> found;
  std::vector<int32_t> starts;
  int64_t nzq = 0;
  for (int64_t j = 0; j < rows; j++) {
    if (std::binary_search(starts.begin(), starts.end(), first[j])) continue;
    if (first[j] >= cols) continue;                      // empty row: not a block
    starts.push_back(first[j]);
    const int32_t h = height[first[j]];
    found.push_back({(int32_t)j, first[j], h, width[first[j]]});
    nzq += (int64_t)h * h;
  }
  if (merge) merge_blocks(found, maxColStep, suggested_block_cols);
  if (nonzero_q_estimate) *nonzero_q_estimate = nzq;
  return emit(found, blocks, capacity, num_blocks);
}

int qrk_detect_blocks(int64_t rows, int64_t cols, const int32_t* csr_outer, const int32_t* csr_inner, int32_t suggested_block_cols,
                      int32_t* blocks, int64_t capacity, int64_t* num_blocks, int64_t* nonzero_q_estimate) {
  return detect_impl(rows, cols, csr_outer, csr_inner, suggested_block_cols, true, blocks, capacity, num_blocks, nonzero_q_estimate);
}

int qrk_detect_band_starts(int64_t rows, int64_t cols, const int32_t* csr_outer, const int32_t* csr_inner, int32_t* blocks, int64_t capacity,
                           int64_t* num_blocks) {
  return detect_impl(rows, cols, csr_outer, csr_inner, 0, false, blocks, capacity, num_blocks, nullptr);
}

int qrk_block_diagonal_pattern(int64_t rows, int64_t cols, int32_t block_rows, int32_t block_cols, int32_t* blocks, int64_t capacity,
                               int64_t* num_blocks) {
  (void)rows;
  if (block_rows <= 0 || block_cols <= 0 || cols < 0) return QRK_STATUS_INVALID_ARGUMENT;
  std::vector<Block> v;
  const int64_t nb = cols / block_cols;
  for (int64_t i = 0; i < nb; i++) v.push_back({(int32_t)(i * block_rows), (int32_t)(i * block_cols), block_rows, block_cols});
  return emit(v, blocks, capacity, num_blocks);
}

int qrk_block_banded_pattern(int64_t rows, int64_t cols, int32_t block_rows, int32_t block_cols, int32_t block_overlap,
                             int32_t suggested_block_cols, int32_t* blocks, int64_t capacity, int64_t* num_blocks) {
  (void)rows;
  const int32_t step = block_cols - block_overlap;
  if (block_rows <= 0 || block_cols <= 0 || step <= 0 || cols < 0) return QRK_STATUS_INVALID_ARGUMENT;
  std::vector<Block> v;
  const int64_t nb = cols / step;
  for (int64_t i = 0; i < nb; i++)      // the last block ends at the matrix bound: block_cols - overlap columns
    v.push_back({(int32_t)(i * block_rows), (int32_t)(i * step), block_rows, (i < nb - 1) ? block_cols : block_cols - block_overlap});
  merge_blocks(v, step, suggested_block_cols);
  return emit(v, blocks, capacity, num_blocks);
}

int qrk_extract_blocks(int64_t rows, int64_t cols, const int32_t* csc_outer, const int32_t* csc_inner, const double* csc_values,
                       const int32_t* row_perm, const int32_t* blocks, int64_t num_blocks, double* values_out) {
  if (rows < 0 || cols < 0 || !csc_outer || !blocks || !values_out || num_blocks < 0) return QRK_STATUS_INVALID_ARGUMENT;
  int64_t off = 0;
  for (int64_t k = 0; k < num_blocks; k++) {
    const int32_t r0 = blocks[4 * k], c0 = blocks[4 * k + 1], nr = blocks[4 * k + 2], nc = blocks[4 * k + 3];
    if (r0 < 0 || c0 < 0 || nr < 0 || nc < 0 || c0 + nc > cols) return QRK_STATUS_INVALID_ARGUMENT;
    std::fill(values_out + off, values_out + off + (int64_t)nr * nc, 0.0);
    for (int32_t j = 0; j < nc; j++) {
      for (int32_t p = csc_outer[c0 + j]; p < csc_outer[c0 + j + 1]; p++) {
        const int32_t i = row_perm ? row_perm[csc_inner[p]] : csc_inner[p];     // row of P * A
        if (i >= r0 && i < r0 + nr) values_out[off + (int64_t)j * nr + (i - r0)] = csc_values[p];
      }
    }
    off += (int64_t)nr * nc;
  }
  return QRK_STATUS_OK;
}

}  // extern "C"
