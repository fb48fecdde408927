// banded_dispatch.hpp — run-time dispatch to the block-banded kernels instantiated per (block_rows,
// block_cols, overlap) in banded_inst.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

namespace qrk {

// ---- the general window chain (banded_generic.cuh) ----
struct GenWindow {            // one entry of the device-side window table
  int row0, col0, nrows, ncols;   // idxRow, idxCol, numRows, numCols (numCols already widened to hold the carried columns)
  int carry;                      // rows carried in from the previous window
  int solved;                     // rows of R finalised by this window
  int steps;                      // Householder steps = min(carry + nrows, ncols)
  int pad_;
  long long voff;                 // offset of the window's dense block in the input values (column-major nrows x ncols_in)
  int ncols_in;                   // columns of the input block (<= ncols)
  int pad2_;
  long long poff;                 // offset of the packed window (column-major (carry + nrows) x ncols) in gen_packed
  long long toff;                 // offset of its tau in gen_tau
  long long coff;                 // offset of its annihilated rows in the complement
};

struct GenArgs {
  const GenWindow* win = nullptr;
  int nwin = 0;
  int max_rows = 0, max_cols = 0;     // largest window (rows = carry + nrows)
  const double* A_in = nullptr;       // dense blocks, back to back
  double* packed = nullptr;           // packed windows: R on / above the diagonal, essential parts below
  double* tau = nullptr;
  const double* b = nullptr;          // right-hand side (n_rows) or vector to transform
  double* y = nullptr;                // thin part (n_cols)
  double* comp = nullptr;             // complement (n_rows - n_cols) or nullptr
  double* x = nullptr;
  long long n_rows = 0, n_cols = 0;
};

struct BandedArgs {
  long long nb = 0;
  int last_cols = 0;          // columns of the last slab that exist in the matrix (1..block_cols)
  const double* A_in = nullptr;
  double* packed = nullptr;
  double* tau = nullptr;
  double* rband = nullptr;
  const double* b = nullptr;
  double* y = nullptr;        // thin part of Q^T b (n_cols)
  double* x = nullptr;
  // two-phase factorisation (banded.cuh): groups of `group` slabs are reduced in parallel, then chased sequentially
  int group = 1;              // slabs per group
  double* gband = nullptr;    // group triangles: groups x W x block_cols, W = (group-1) S + block_cols
  double* gy = nullptr;       // groups x W: pivot-row values of the right-hand side between the two phases
  double* cvec = nullptr;     // chase reflectors: groups x W x overlap raw tails
  double* ctau = nullptr;     // groups x W x {tau, inv}: v = [1; inv * raw tail]
  double* comp = nullptr;     // apply_qt only, optional: the complement of Q^T b (banded.cuh, banded_apply_qt_kernel); see banded_comp_rows
  // apply_qt on several columns at once (ncols > 1): leading dimensions of b, y (thin part) and comp; gy must hold ncols * groups * W
  int ncols = 1;
  long long ldb = 0, ldy = 0, ldcomp = 0;
  const GenArgs* gen = nullptr;   // set for handles on the general window chain: the vtable below then runs banded_generic.cuh
  long long ldx = 0;              // apply_q on several columns (general chain only)
};

struct BandedVTable {
  int br, bc, ov;
  cudaError_t (*factor)(const BandedArgs&, cudaStream_t);
  cudaError_t (*apply_qt)(const BandedArgs&, cudaStream_t);   // b (n_rows) -> y (thin part, n_cols)
  cudaError_t (*apply_q)(const BandedArgs&, cudaStream_t);    // y (thin part, n_cols; zero complement) -> x = Q1 y (n_rows)
  cudaError_t (*backsolve)(const BandedArgs&, cudaStream_t);
};

const BandedVTable* banded_vtable(int br, int bc, int ov);
const BandedVTable* banded_generic_vtable();           // any window chain (BandedArgs::gen), one CTA, sequential
size_t banded_generic_smem_bytes(int max_rows, int max_cols);
cudaError_t banded_generic_export_r(const GenArgs& g, const int* d_col0, const int* d_outer, const int* d_inner, double* d_vals, cudaStream_t s);
// structure.cpp: the reference's merged windows for a slab geometry, {idxRow, idxCol, numRows, numCols} each
void banded_reference_windows(long long nb, int br, int bc, int ov, int last_cols, int suggested, std::vector<int32_t>& out4);
void banded_reference_windows_from_blocks(const std::vector<int32_t>& in4, int suggested, std::vector<int32_t>& out4);
int banded_launches_per_call();
// rows of the complement output of apply_qt: nb (OV + BR - BC) window rows + groups * OV chase rows + (BC - last_cols)
long long banded_comp_rows(long long nb, int br, int bc, int ov, int group, int last_cols);   // kernels per factor / apply call (for the launch counter)   // nullptr when the shape is not instantiated

}  // namespace qrk
