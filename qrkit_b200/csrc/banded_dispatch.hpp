// banded_dispatch.hpp — run-time dispatch to the block-banded kernels instantiated per (block_rows,
// block_cols, overlap) in banded_inst.cu.
#pragma once
#include <cuda_runtime.h>

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
  double* ycomp = nullptr;    // annihilated rows' part of Q^T b (n_rows - n_cols), optional
  double* x = nullptr;
};

struct BandedVTable {
  int br, bc, ov;
  cudaError_t (*factor)(const BandedArgs&, cudaStream_t);
  cudaError_t (*apply_qt)(const BandedArgs&, cudaStream_t);
  cudaError_t (*backsolve)(const BandedArgs&, cudaStream_t);
};

const BandedVTable* banded_vtable(int br, int bc, int ov);   // nullptr when the shape is not instantiated

}  // namespace qrk
