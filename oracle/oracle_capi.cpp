// oracle_capi.cpp — C entry points of the CPU ORACLE for ctypes (tests / smoke / bench cpu_baseline only).
// TEST INFRASTRUCTURE: the product (qrkit_b200/) never links or loads this library.
// See qrkit_oracle.hpp for what is restated and the "PARITY UNPINNED" statement.
#include "qrkit_oracle.hpp"

#include <chrono>
#include <thread>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace qrk_oracle;

namespace {
BlockDiag make_bd(int nb, const int* br, const int* bc, const double* values, int nRows, int nCols) {
  BlockDiag m;
  m.nRows = nRows; m.nCols = nCols;
  m.br.assign(br, br + nb); m.bc.assign(bc, bc + nb);
  m.off.resize(nb);
  int64_t o = 0;
  for (int i = 0; i < nb; i++) { m.off[i] = o; o += (int64_t)br[i] * bc[i]; }
  m.values.assign(values, values + o);
  return m;
}
void copy_sparse(const Sparse& S, int* outer, int* inner, double* val) {
  std::copy(S.outer.begin(), S.outer.end(), outer);
  std::copy(S.inner.begin(), S.inner.end(), inner);
  std::copy(S.val.begin(), S.val.end(), val);
}
Sparse make_csc(int rows, int cols, const int* outer, const int* inner, const double* val) {
  Sparse S; S.row_major = false; S.rows = rows; S.cols = cols;
  S.outer.assign(outer, outer + cols + 1);
  S.inner.assign(inner, inner + outer[cols]);
  S.val.assign(val, val + outer[cols]);
  return S;
}
struct BDHandle { BlockDiag mat; BlockDiagQR f; };
struct BandedHandle { BandedQR f; std::vector<BlockInfo> blocks; };
struct AngularHandle { AngularQR f; };
}  // namespace

extern "C" {

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }
int orc_omp_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---- dense kernels -----------------------------------------------------------------------
void orc_householder_qr(double* A, int r, int c, double* tau) { householder_qr(A, r, c, tau); }
int orc_colpiv_qr(double* A, int r, int c, double* tau, int* perm) { return colpiv_householder_qr(A, r, c, tau, perm); }
void orc_householder_q(const double* QR, int r, int c, const double* tau, double* Q) {
  Dense q = householder_q(QR, r, c, tau);
  std::copy(q.v.begin(), q.v.end(), Q);
}
void orc_block_t_factor(const double* V, int rows, int n, const double* tau, double* T) {
  Dense t = block_householder_t_factor(V, rows, rows, n, tau);
  std::copy(t.v.begin(), t.v.end(), T);
}
double orc_synth_value(uint64_t seed, uint64_t block, uint64_t row, uint64_t col, double lo, double hi) {
  return synth_value(seed, block, row, col, lo, hi);
}
// values for nb uniform r x c blocks, block-major / col-major inside the block
void orc_synth_fill_blocks(uint64_t seed, int64_t block0, int64_t nb, int r, int c, double lo, double hi, double* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nb; i++)
    for (int j = 0; j < c; j++)
      for (int k = 0; k < r; k++) out[(size_t)i * r * c + (size_t)j * r + k] = synth_value(seed, block0 + i, k, j, lo, hi);
}

// ---- block diagonal (reference-faithful) ---------------------------------------------------
void* orc_bd_factorize(int nb, const int* br, const int* bc, const double* values, int nRows, int nCols, int colpiv,
                       int qformat, int build_q) {
  auto* h = new BDHandle;
  h->mat = make_bd(nb, br, bc, values, nRows, nCols);
  h->f = block_diagonal_factorize(h->mat, colpiv != 0, qformat, build_q != 0);
  return h;
}
void orc_bd_free(void* hv) { delete (BDHandle*)hv; }
int orc_bd_info(void* hv) { return ((BDHandle*)hv)->f.info; }
int orc_bd_rank(void* hv) { return ((BDHandle*)hv)->f.rank; }
int64_t orc_bd_q_nnz(void* hv) { return (int64_t)((BDHandle*)hv)->f.Q.nnz(); }
int64_t orc_bd_r_nnz(void* hv) { return (int64_t)((BDHandle*)hv)->f.R.nnz(); }
void orc_bd_get_q(void* hv, int* outer, int* inner, double* val) { copy_sparse(((BDHandle*)hv)->f.Q, outer, inner, val); }
void orc_bd_get_r(void* hv, int* outer, int* inner, double* val) { copy_sparse(((BDHandle*)hv)->f.R, outer, inner, val); }
void orc_bd_get_perms(void* hv, int* colperm, int* rowperm) {
  auto* h = (BDHandle*)hv;
  std::copy(h->f.colPerm.begin(), h->f.colPerm.end(), colperm);
  if (rowperm) std::copy(h->f.rowPerm.begin(), h->f.rowPerm.end(), rowperm);
}
void orc_bd_get_packed(void* hv, double* packed, double* tau) {
  auto* h = (BDHandle*)hv;
  std::copy(h->f.packed.begin(), h->f.packed.end(), packed);
  std::copy(h->f.tau.begin(), h->f.tau.end(), tau);
}
void orc_bd_solve(void* hv, const double* b, double* x) {
  auto* h = (BDHandle*)hv;
  std::vector<double> xv = block_diagonal_solve(h->f, h->mat.nRows, h->mat.nCols, b);
  std::copy(xv.begin(), xv.end(), x);
}
void orc_bd_apply_qt(void* hv, const double* b, double* y) {  // y = matrixQ().transpose() * b
  auto* h = (BDHandle*)hv;
  spmv_transposed(h->f.Q, b, h->mat.nRows, y, h->mat.nRows, 1);
}
void orc_bd_apply_q(void* hv, const double* b, double* y) {
  auto* h = (BDHandle*)hv;
  spmv(h->f.Q, b, h->mat.nRows, y, h->mat.nRows, 1);
}

// ---- block diagonal, compact variants used as the CPU baseline (uniform blocks) ---------------
// Variant B (threads = 1) / C (threads > 1): packed Householder vectors + fused Q^T b + back substitution.
// values is overwritten with the packed factors.  Returns seconds.
double orc_bd_compact_uniform(int64_t nb, int r, int c, double* values, const double* b, double* x, double* tau,
                              int* perm, int colpiv, int threads) {
  auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
  for (int64_t i = 0; i < nb; i++) {
    double* A = values + (size_t)i * r * c;
    double* ta = tau + (size_t)i * c;
    int pstack[64];
    std::vector<int> pheap(c > 64 ? c : 0);
    int* p = c > 64 ? pheap.data() : pstack;
    if (colpiv) colpiv_householder_qr(A, r, c, ta, p);
    else { householder_qr(A, r, c, ta); for (int j = 0; j < c; j++) p[j] = j; }
    if (perm) for (int j = 0; j < c; j++) perm[(size_t)i * c + j] = (int)(i * c + p[j]);
    if (b) {
      double bstack[256];
      std::vector<double> bheap(r > 256 ? r : 0);
      double* bb = r > 256 ? bheap.data() : bstack;
      for (int k = 0; k < r; k++) bb[k] = b[(size_t)i * r + k];
      apply_qt_inplace(A, r, c, ta, bb, r, 1);
      for (int j = c - 1; j >= 0; --j) {
        double s = bb[j];
        for (int k = j + 1; k < c; k++) s -= A[(size_t)k * r + j] * bb[k];
        bb[j] = s / A[(size_t)j * r + j];
      }
      for (int j = 0; j < c; j++) x[(size_t)i * c + p[j]] = bb[j];
    }
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Variant A: the reference's algorithm and data structures (explicit Q, sparse assembly, sparse Q^T b,
// sparse triangular solve).  Returns seconds for factorize + solve.
double orc_bd_reference_uniform(int64_t nb, int r, int c, const double* values, const double* b, double* x, int colpiv) {
  BlockDiag m;
  m.nRows = (int)(nb * r); m.nCols = (int)(nb * c);
  m.br.assign(nb, r); m.bc.assign(nb, c); m.off.resize(nb);
  for (int64_t i = 0; i < nb; i++) m.off[i] = i * r * c;
  m.values.assign(values, values + (size_t)nb * r * c);
  auto t0 = std::chrono::steady_clock::now();
  BlockDiagQR f = block_diagonal_factorize(m, colpiv != 0, FullQ, true);
  std::vector<double> xv = block_diagonal_solve(f, m.nRows, m.nCols, b);
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::copy(xv.begin(), xv.end(), x);
  return dt;
}

// ---- block structure --------------------------------------------------------------------------
int orc_block_banded_pattern(int matRows, int matCols, int blockRows, int blockCols, int overlap, int suggested,
                             int* out /* 4 ints per block */, int cap) {
  auto v = from_block_banded_pattern(matRows, matCols, blockRows, blockCols, overlap, suggested);
  for (int i = 0; i < (int)v.size() && i < cap; i++) {
    out[4 * i] = v[i].idxRow; out[4 * i + 1] = v[i].idxCol; out[4 * i + 2] = v[i].numRows; out[4 * i + 3] = v[i].numCols;
  }
  return (int)v.size();
}

// ---- banded --------------------------------------------------------------------------------------
void* orc_banded_factorize(int rows, int cols, const int* outer, const int* inner, const double* val, int nblocks,
                           const int* blocks4) {
  auto* h = new BandedHandle;
  Sparse A = make_csc(rows, cols, outer, inner, val);
  for (int i = 0; i < nblocks; i++) h->blocks.push_back(BlockInfo{blocks4[4 * i], blocks4[4 * i + 1], blocks4[4 * i + 2], blocks4[4 * i + 3]});
  h->f = banded_factorize(A, h->blocks);
  return h;
}
void orc_banded_free(void* hv) { delete (BandedHandle*)hv; }
int64_t orc_banded_r_nnz(void* hv) { return (int64_t)((BandedHandle*)hv)->f.R.nnz(); }
void orc_banded_get_r(void* hv, int* outer, int* inner, double* val) { copy_sparse(((BandedHandle*)hv)->f.R, outer, inner, val); }
int orc_banded_num_blocks(void* hv) { return (int)((BandedHandle*)hv)->f.blocks.size(); }
void orc_banded_block_dims(void* hv, int k, int* out5) {  // rows, cols, row, col, numZeros
  const YTYBlock& b = ((BandedHandle*)hv)->f.blocks[k];
  out5[0] = b.Y.rows; out5[1] = b.Y.cols; out5[2] = b.row; out5[3] = b.col; out5[4] = b.numZeros;
}
void orc_banded_get_block(void* hv, int k, double* Y, double* T) {
  const YTYBlock& b = ((BandedHandle*)hv)->f.blocks[k];
  std::copy(b.Y.v.begin(), b.Y.v.end(), Y);
  std::copy(b.T.v.begin(), b.T.v.end(), T);
}
void orc_banded_apply_q(void* hv, double* v, int transpose) { ytysequence_apply(((BandedHandle*)hv)->f.blocks, v, transpose != 0); }
void orc_banded_solve(void* hv, const double* b, double* x) {
  auto xv = banded_solve(((BandedHandle*)hv)->f, b);
  std::copy(xv.begin(), xv.end(), x);
}

// ---- block angular --------------------------------------------------------------------------------
// left = block diagonal
void* orc_angular_factorize_bd(int nb, const int* br, const int* bc, const double* values, int nRows, int nCols,
                               const double* J2 /* nRows x m2 col-major */, int m2, int left_colpiv, int right_kind, int panel) {
  auto* h = new AngularHandle;
  BlockDiag m = make_bd(nb, br, bc, values, nRows, nCols);
  Dense D(nRows, m2);
  std::copy(J2, J2 + (size_t)nRows * m2, D.v.begin());
  h->f = angular_factorize_bd(m, D, left_colpiv != 0, right_kind, panel);
  return h;
}
// left = banded
void* orc_angular_factorize_banded(int rows, int cols, const int* outer, const int* inner, const double* val, int nblocks,
                                   const int* blocks4, const double* J2, int m2, int right_kind, int panel) {
  auto* h = new AngularHandle;
  Sparse A = make_csc(rows, cols, outer, inner, val);
  std::vector<BlockInfo> bl;
  for (int i = 0; i < nblocks; i++) bl.push_back(BlockInfo{blocks4[4 * i], blocks4[4 * i + 1], blocks4[4 * i + 2], blocks4[4 * i + 3]});
  Dense D(rows, m2);
  std::copy(J2, J2 + (size_t)rows * m2, D.v.begin());
  h->f = angular_factorize_banded(A, bl, D, right_kind, panel);
  return h;
}
void orc_angular_free(void* hv) { delete (AngularHandle*)hv; }
int orc_angular_rank(void* hv) { return ((AngularHandle*)hv)->f.rank; }
int64_t orc_angular_r_nnz(void* hv) { return (int64_t)((AngularHandle*)hv)->f.R.nnz(); }
void orc_angular_get_r(void* hv, int* outer, int* inner, double* val) { copy_sparse(((AngularHandle*)hv)->f.R, outer, inner, val); }
void orc_angular_get_perms(void* hv, int* colperm, int* rowperm) {
  auto* h = (AngularHandle*)hv;
  std::copy(h->f.colPerm.begin(), h->f.colPerm.end(), colperm);
  if (rowperm) std::copy(h->f.rowPerm.begin(), h->f.rowPerm.end(), rowperm);
}
void orc_angular_get_right_perm(void* hv, int* p) {
  auto* h = (AngularHandle*)hv;
  std::copy(h->f.right.perm.begin(), h->f.right.perm.end(), p);
}
void orc_angular_apply_qt(void* hv, const double* v, double* y) {
  auto* h = (AngularHandle*)hv;
  auto r = angular_apply_qt(h->f, v);
  std::copy(r.begin(), r.end(), y);
}
void orc_angular_solve(void* hv, const double* b, double* x) {
  auto* h = (AngularHandle*)hv;
  auto r = angular_solve(h->f, b);
  std::copy(r.begin(), r.end(), x);
}
// timing leg for the block-angular CPU baseline (left = uniform block diagonal r x c, dense ColPiv border).
double orc_angular_reference_uniform(int64_t nb, int r, int c, const double* values, const double* J2, int m2,
                                     const double* b, double* x, int left_colpiv) {
  BlockDiag m;
  m.nRows = (int)(nb * r); m.nCols = (int)(nb * c);
  m.br.assign(nb, r); m.bc.assign(nb, c); m.off.resize(nb);
  for (int64_t i = 0; i < nb; i++) m.off[i] = i * r * c;
  m.values.assign(values, values + (size_t)nb * r * c);
  Dense D(m.nRows, m2);
  std::copy(J2, J2 + (size_t)m.nRows * m2, D.v.begin());
  auto t0 = std::chrono::steady_clock::now();
  AngularQR f = angular_factorize_bd(m, D, left_colpiv != 0, 0, 2);
  auto xv = angular_solve(f, b);
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::copy(xv.begin(), xv.end(), x);
  return dt;
}

}  // extern "C"
