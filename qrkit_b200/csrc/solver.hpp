// solver.hpp — internal state behind the opaque qrk_handle_t (host side of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/qrkit_b200.h"
#include "angular_dispatch.hpp"
#include "banded_dispatch.hpp"

namespace qrk {

struct SizeClass {          // one launch of the generic kernel: blocks with a similar shared-memory need
  int warps = 1;            // team size W
  int wy_mr = 0;            // 0: bd_generic kernel; 1/2/4: bd_wy kernel with that many panel rows per lane
  size_t smem = 0;          // dynamic shared memory per CTA
  long long count = 0;
  int* d_ids = nullptr;     // block numbers of this class (nullptr: all blocks, identity order)
};

}  // namespace qrk

struct qrk_solver {
  qrk_desc_t desc{};
  int device = 0;
  cudaStream_t stream = nullptr;      // the stream work is enqueued on
  cudaStream_t own_stream = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;     // copy streams of the chunked host pipeline (qrk_compute_solve, QRK_HOST)
  std::vector<cudaEvent_t> pipe_events;
  cudaStream_t s_aux = nullptr;                     // look-ahead stream of the blocked dense border (panel p+1 beside update p)
  std::vector<cudaEvent_t> aux_events;
  std::string err;
  long long launches = 0;

  // ---- block structure (SparseBlockDiagonal / block-COO index) ----
  long long nb = 0;
  bool uniform = true;
  int ur = 0, uc = 0, max_r = 0, max_c = 0;
  std::vector<int> h_rows, h_cols;
  std::vector<long long> h_voff, h_roff, h_coff;
  int *d_rows = nullptr, *d_cols = nullptr;
  long long *d_voff = nullptr, *d_roff = nullptr, *d_coff = nullptr;
  long long n_rows = 0, n_cols = 0, sum_rows = 0, sum_cols = 0, total_values = 0;
  bool small_path = false;            // thread-per-block kernels instantiated for (ur, uc)
  std::vector<qrk::SizeClass> classes;

  // ---- factorisation state ----
  double* d_values = nullptr;         // blocks, overwritten by the packed factors
  bool own_values = false;
  double* d_tau = nullptr;
  int* d_perm = nullptr;              // colsPermutation().indices(), int32[n_cols]
  std::vector<int> h_rowperm;         // rowsPermutation().indices()
  bool analyzed = false, has_blocks = false, factorized = false;
  int info = QRK_INFO_SUCCESS;

  // ---- block angular (kind == QRK_BLOCK_ANGULAR): dense border of m2 columns ----
  int m2 = 0;
  const qrk::AngularVTable* avt = nullptr;
  int a_grid = 0;                      // CTAs of the factor kernel = partial triangles
  int world = 1;                       // > 1: compute stops at the per-GPU triangle, qrk_angular_merge finishes
  const double* d_border = nullptr;    // n x m2 column-major (borrowed device pointer or d_border_own)
  long long ld_border = 0;
  double* d_border_own = nullptr;
  size_t cap_border = 0;
  double *d_atop = nullptr, *d_y1 = nullptr, *d_abot = nullptr, *d_partials = nullptr, *d_tri = nullptr, *d_root = nullptr;
  int* d_root_i = nullptr;
  bool have_abot = false;              // the residual panel of the last compute() is resident (solve(b) possible)
  bool root_done = false;
  bool pending = false;                // a right-hand side waits for qrk_angular_merge
  // fused peer exchange over NVLink (qrk_angular_p2p_attach): the triangles travel inside the root kernel
  double* d_xchg = nullptr;            // this rank's exchange buffer: [2][world][Tri] doubles + [2][world] uint64 flags
  size_t xchg_bytes = 0;
  double** d_xchg_peers = nullptr;     // device array of world pointers (peer buffers as mapped in this process)
  int* d_xchg_err = nullptr;
  unsigned long long* d_xchg_seq = nullptr;   // step counter, advanced by the root kernel itself
  int xchg_rank = -1;                  // >= 0 once attached
  unsigned long long xchg_timeout_ns = 0;   // 0: default
  // Q2 of the fused TSQR path, built on demand for matrixQ() products (the TSQR tree keeps no reflectors): Householder QR of
  // the kept residual panel in the column order P2, plus the row signs that align it with the stored R2
  double *d_q2 = nullptr, *d_q2tau = nullptr, *d_q2sign = nullptr, *d_q2scr = nullptr, *d_qtmp = nullptr;
  int* d_q2iscr = nullptr;
  size_t cap_qtmp = 0;
  double* d_qthin = nullptr;           // scratch of the thin-factor products (qrk_apply_qt_thin / qrk_apply_q_thin)
  size_t cap_qthin = 0;
  bool q2_ready = false;
  double* pending_x = nullptr;
  int pending_space = 0;
  int pending_keep_rhs_only = 0;

  // ---- block angular, wide border or a left block outside the TSQR kernels (dense_border.cuh) ----
  bool wide = false;
  double* d_wx = nullptr;              // n x (m2 + 1): Q1^T [J2 | b]; rows [m1, n) hold the right block's packed QR
  double *d_wupd = nullptr, *d_wdir = nullptr, *d_wtau2 = nullptr, *d_wscal = nullptr;
  // blocked first stage of the dense right block (dense_blocked.cuh): its tau, the per-panel T factors, and the
  // M x (M + 1) triangle the ColPiv second stage works on; wide_blocked: the last compute() took that path
  double *d_wtau1 = nullptr, *d_wT = nullptr, *d_wtri = nullptr, *d_wpart = nullptr;   // d_wpart: partial W = V^T A2 per (column block, row range)
  bool wide_blocked = false;
  // a launch-bound step replayed from a CUDA graph once the same buffers come a second time (capi.cu: step_graph_run):
  // the wide-border step (~250 launches on two streams) and the three-launch TSQR step of the narrow border
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    const void* key[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int seen = 0;
    bool off = false, flag = false;
    long long launches = 0;
  };
  StepGraph sg_wide, sg_tsqr, sg_solve;
  bool thin_deferred = false;          // QRK_RIGHT_THIN_SPARSE: the last factorisation deferred zero-pivot columns (stored-factor products refused)
  // dense right-block path: leading dimension of d_wx (rows of [thin part ; complement] of Q1^T [J2 | b]) and the number of
  // complement rows the right block is factored on.  Block-diagonal left: n_rows and n_rows - m1; banded left: m1 + the
  // extended complement of the two-phase application (banded_comp_rows)
  long long w_ld = 0, w_N = 0;
  bool left_banded = false;            // BlockAngularSparseQR<BandedBlockedSparseQR, ...>
  int *d_wperm = nullptr, *d_wiscal = nullptr;

  // ---- banded blocked (kind == QRK_BANDED_BLOCKED): sequential window sweep on one SM ----
  const qrk::BandedVTable* bvt = nullptr;
  int b_ov = 0, b_step = 0;            // overlap, column step S = block_cols - overlap
  double *d_rband = nullptr;           // band R: n_cols x block_cols
  double *d_btau = nullptr;            // tau: num_blocks x block_cols
  double *d_ythin = nullptr;           // (Q^T b)[0:n_cols]
  std::vector<int32_t> b_windows;      // the reference's merged windows {idxRow, idxCol, numRows, numCols}: the stored pattern of matrixR()
  int b_group = 1;                     // slabs per parallel group of the two-phase banded factorisation (banded.cuh)
  double *d_gband = nullptr, *d_gy = nullptr, *d_cvec = nullptr, *d_ctau = nullptr;   // group triangles / chase reflectors
  // the general window chain (banded_generic.cuh): block sizes / overlaps at run time, an exact n x n Q
  bool bgen = false;
  std::vector<qrk::GenWindow> g_win;
  qrk::GenArgs g_args;                 // device pointers of the chain (win, packed, tau), sizes
  qrk::GenWindow* d_gwin = nullptr;
  int* d_gcol0 = nullptr;
  double *d_gpacked = nullptr, *d_gtau = nullptr, *d_gcomp = nullptr;
  size_t cap_gy = 0;                   // doubles in d_gy (one vector of groups x W per right-hand side column)

  // ---- staging buffers for host-memspace calls ----
  double *d_b = nullptr, *d_x = nullptr;
  size_t cap_b = 0, cap_x = 0;
};
