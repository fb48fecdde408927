// banded_dispatch.hpp — run-time dispatch to the block-banded kernels instantiated per (block_rows,
// block_cols, overlap) in banded_inst.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

namespace qrk {

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
};

struct BandedVTable {
  int br, bc, ov;
  cudaError_t (*factor)(const BandedArgs&, cudaStream_t);
  cudaError_t (*apply_qt)(const BandedArgs&, cudaStream_t);   // b (n_rows) -> y (thin part, n_cols)
  cudaError_t (*apply_q)(const BandedArgs&, cudaStream_t);    // y (thin part, n_cols; zero complement) -> x = Q1 y (n_rows)
  cudaError_t (*backsolve)(const BandedArgs&, cudaStream_t);
};

const BandedVTable* banded_vtable(int br, int bc, int ov);
// structure.cpp: the reference's merged windows for a slab geometry, {idxRow, idxCol, numRows, numCols} each
void banded_reference_windows(long long nb, int br, int bc, int ov, int last_cols, int suggested, std::vector<int32_t>& out4);
int banded_launches_per_call();
// rows of the complement output of apply_qt: nb (OV + BR - BC) window rows + groups * OV chase rows + (BC - last_cols)
long long banded_comp_rows(long long nb, int br, int bc, int ov, int group, int last_cols);   // kernels per factor / apply call (for the launch counter)   // nullptr when the shape is not instantiated

}  // namespace qrk
