/* qrkit_b200.h — C ABI of the B200-native structured-sparse QR hot path (drop-in boundary).
 *
 * This is the only surface the host side binds: plain pointers and sizes, no C++/torch types.
 * A C++ façade with the reference's Eigen-style method names (include/qrkit_b200/QRKit.hpp) forwards
 * to it; INTEGRATION.md shows the stub a QRKit maintainer would add on the reference side.
 *
 * Each entry point names the reference interface it replaces (file:line relative to the
 * jasvob/QRKit tree).  All functions
 *   - return a qrk_status (0 = OK) and never throw,
 *   - are ordered on the handle's CUDA stream (qrk_set_stream); host-pointer variants return after
 *     the result is in the caller's buffer, device-pointer variants return after enqueueing,
 *   - use one handle from one host thread at a time (as the reference's solver objects,
 *     BlockDiagonalSparseQR.h:316 `mutable m_info`).
 * There is no CPU fallback: without a CUDA device every compute entry point returns
 * QRK_STATUS_NO_DEVICE.
 *
 * Data layout ("block-COO", the device form of SparseBlockCOO.h:37-50 / SparseBlockDiagonal.h:44-163):
 *   values[]  one flat FP64 array, block i at values + val_off(i), column-major r_i x c_i inside
 *             the block (Eigen's Matrix<double,r,c> layout, so a std::vector of fixed-size
 *             blocks IS this array);
 *   block i sits at matrix position (base_row(i), base_col(i)) = prefix sums of the block sizes
 *             (BlockDiagonalSparseQR.h:524-525; SparseQRUtils.h:255-272 for the uniform case).
 * All indices are int32 (StorageIndex = int, test/test-qrkit.cpp:40), all values FP64.
 */
#ifndef QRKIT_B200_H_
#define QRKIT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define QRK_API
#else
#define QRK_API __attribute__((visibility("default")))
#endif

typedef struct qrk_solver* qrk_handle_t;

typedef enum {
  QRK_STATUS_OK = 0,
  QRK_STATUS_INVALID_ARGUMENT = 1,
  QRK_STATUS_NOT_FACTORIZED = 2,   /* eigen_assert "The factorization should be called first" (BlockDiagonalSparseQR.h:260) */
  QRK_STATUS_CUDA_ERROR = 3,
  QRK_STATUS_NO_DEVICE = 4,
  QRK_STATUS_ALLOC_FAILED = 5,
  QRK_STATUS_UNSUPPORTED = 6,
  QRK_STATUS_PEER_TIMEOUT = 7      /* fused peer exchange: a peer GPU never delivered its triangle; results are NaN-poisoned */
} qrk_status;

/* Eigen::ComputationInfo values, as returned by info() (BlockDiagonalSparseQR.h:309-313). */
typedef enum { QRK_INFO_SUCCESS = 0, QRK_INFO_NUMERICAL_ISSUE = 1, QRK_INFO_NO_CONVERGENCE = 2, QRK_INFO_INVALID_INPUT = 3 } qrk_computation_info;

/* Which reference solver the handle stands for. */
typedef enum {
  QRK_BLOCK_DIAGONAL = 0,  /* BlockDiagonalSparseQR  (BlockDiagonalSparseQR.h:37-335)  */
  QRK_BLOCK_ANGULAR = 1,   /* BlockAngularSparseQR   (BlockAngularSparseQR.h:79-281), left = block diagonal */
  QRK_BANDED_BLOCKED = 2   /* BandedBlockedSparseQR  (BandedBlockedSparseQR.h:122-344) */
} qrk_kind;

/* The per-block dense solver (template parameter _BlockQRSolver, BlockDiagonalSparseQR.h:37). */
typedef enum {
  QRK_PIVOT_NONE = 0,   /* Eigen::HouseholderQR        */
  QRK_PIVOT_COLPIV = 1  /* Eigen::ColPivHouseholderQR (test/test-qrkit.cpp:49-51) */
} qrk_pivoting;

/* The dense right-block solver of BlockAngularSparseQR (template parameter RightSolver, BlockAngularSparseQR.h:79). */
typedef enum {
  QRK_RIGHT_COLPIV = 0,    /* Eigen::ColPivHouseholderQR<MatrixXd> (test/test-qrkit.cpp:46-48, examples/ellipse_fitting.cpp:35) */
  QRK_RIGHT_UNPIVOTED = 1, /* BlockedThinDenseQR<MatrixXd, N> / Eigen::HouseholderQR: no column pivoting, P2 = identity,
                              rank() = cols (BlockedThinDenseQR.h:132); R2 equals the reference's up to row signs */
  QRK_RIGHT_THIN_SPARSE = 2 /* BlockedThinSparseQR<SparseMatrix, SuggestedBlockCols> on the (dense) border (BlockedThinSparseQR.h:105-166,
                              198-283; test/test-qrkit.cpp:54-57, 329-362): ColPivHouseholderQR inside every panel of
                              SuggestedBlockCols columns (desc.reserved[1], 0 = the reference's 2) with Eigen's per-panel
                              nonzero-pivot rule; columns whose pivot is "zero" are deferred to the end of P2 (:250-255, 150-158);
                              rank() = the nonzero pivots (:281).  With full column rank R2, P2 and x equal the reference's.
                              Rank-deficient borders: the reference's R collides (its R columns restart at m_nonzeroPivots + bc);
                              here the deferred columns receive every later reflector and take the trailing columns of R, so
                              A P = Q R holds and x is the basic solution; solve() / matrixQ() products on a STORED factorisation
                              that deferred columns are refused (use the fused qrk_compute_solve) */
} qrk_right_solver;

/* The left solver of BlockAngularSparseQR (template parameter LeftSolver, BlockAngularSparseQR.h:79). */
typedef enum {
  QRK_LEFT_BLOCK_DIAGONAL = 0,   /* BlockDiagonalSparseQR (examples/ellipse_fitting.cpp:36-37; the QRkitBD column of README.md:29) */
  QRK_LEFT_BANDED_BLOCKED = 1    /* BandedBlockedSparseQR (test/test-qrkit.cpp:44-48; the QRkitBB column): J1 is given as num_blocks
                                    slabs of block_rows x block_cols with block_overlap, as for kind QRK_BANDED_BLOCKED */
} qrk_left_solver;

/* MatrixQFormat (BlockDiagonalSparseQR.h:59-62). */
typedef enum { QRK_FULL_Q = 0, QRK_BLOCK_DIAGONAL_Q = 1 } qrk_qformat;

/* Where a caller's buffer lives. */
typedef enum { QRK_HOST = 0, QRK_DEVICE = 1 } qrk_memspace;

typedef struct {
  int32_t kind;            /* qrk_kind */
  int32_t device;          /* CUDA device ordinal */
  int64_t num_blocks;      /* nb */
  int32_t block_rows;      /* uniform block size (fromBlockDiagonalPattern, SparseBlockDiagonal.h:72-89); */
  int32_t block_cols;      /*   both 0 => per-block sizes come from rows[] / cols[] */
  const int32_t* rows;     /* host arrays of length nb (copied), or NULL when uniform */
  const int32_t* cols;
  int64_t n_rows;          /* matrix rows; 0 => sum of block rows.  Rows beyond the blocks get Q(i,i)=1 (BlockDiagonalSparseQR.h:530-533) */
  int64_t n_cols;          /* matrix cols; 0 => sum of block cols */
  int32_t pivoting;        /* qrk_pivoting */
  int32_t q_format;        /* qrk_qformat */
  int32_t border_cols;     /* block angular: m2 = columns of the dense right block (BlockAngularSparseQR.h:464); else 0 */
  int32_t block_overlap;   /* banded: column overlap of consecutive blocks (_BlockOverlap, BandedBlockedSparseQR.h:122); else 0 */
  int32_t right_solver;    /* block angular: qrk_right_solver, the RightSolver template argument (BlockAngularSparseQR.h:79) */
  int32_t left_solver;     /* block angular: qrk_left_solver, the LeftSolver template argument (BlockAngularSparseQR.h:79) */
  int32_t reserved[2];     /* [0]: banded: SuggestedBlockCols template argument of BandedBlockedSparseQR (:122), 0 = its default 2 —
                              it only shapes the windows the reference merges, i.e. the stored pattern of matrixR();
                              [1]: block angular with QRK_RIGHT_THIN_SPARSE: SuggestedBlockCols of BlockedThinSparseQR, 0 = 2 */
} qrk_desc_t;

/* ---- library ------------------------------------------------------------------------------- */
QRK_API int qrk_version(void);
QRK_API const char* qrk_status_string(int status);
QRK_API const char* qrk_last_error(qrk_handle_t h);          /* lastErrorMessage() (BlockAngularSparseQR.h:199) but actually filled */
QRK_API int qrk_device_count(int* count);

/* ---- life cycle ------------------------------------------------------------------------------ */
/* Solver construction (default ctor BlockDiagonalSparseQR.h:74) + the structure half of
 * analyzePattern (:392-405): sizes R, builds the prefix-sum block index on the device. */
QRK_API int qrk_create(const qrk_desc_t* desc, qrk_handle_t* out);
QRK_API int qrk_destroy(qrk_handle_t h);
QRK_API int qrk_set_stream(qrk_handle_t h, void* cuda_stream);   /* NULL => the handle's own stream */
QRK_API int qrk_synchronize(qrk_handle_t h);

/* ---- SparseBlockCOO assembly / upload (SparseBlockDiagonal.h:44-163) --------------------------- */
/* Copy the blocks into the handle's device storage (host: H2D on the handle's stream;
 * device: D2D).  `values` holds total_values doubles in the block-COO layout above. */
QRK_API int qrk_set_blocks(qrk_handle_t h, const double* values, int memspace);
/* Zero-copy: adopt a caller-owned device buffer; factorize() then overwrites it in place with the
 * packed factors (R in the upper triangle, essential Householder parts below). */
QRK_API int qrk_adopt_blocks(qrk_handle_t h, double* device_values);
QRK_API int qrk_total_values(qrk_handle_t h, int64_t* n);

/* Host side of the upload (the reference keeps the blocks in a std::vector<Block> on the host, SparseBlockDiagonal.h:46,159).
 * Host-memspace calls run at PCIe speed only from PINNED memory that is NUMA-local to the GPU; with one process per GPU on a
 * two-socket host, buffers that all land on one socket cap the aggregate upload (measured: 1/2/4/8 ranks 59/57/30/22 GB/s each).
 *   qrk_bind_host_thread_to_device  pins the CALLING thread to the CPUs of the GPU's PCIe root complex (sysfs local_cpulist of
 *                                   its bus id) and makes that NUMA node the preferred one for the thread's allocations;
 *                                   *numa_node = -1 and *cpus_bound = 0 when the platform does not expose the topology
 *   qrk_host_alloc / qrk_host_free  page-locked host memory (cudaHostAlloc), allocated and first-touched under that binding */
QRK_API int qrk_bind_host_thread_to_device(int32_t device, int32_t* numa_node, int32_t* cpus_bound);
QRK_API int qrk_host_alloc(void** ptr, int64_t bytes, int32_t device);
QRK_API int qrk_host_free(void* ptr);

/* ---- compute / factorize (BlockDiagonalSparseQR.h:94-104, 415-547) ----------------------------- */
/* row_perm: optional int32[n_rows] host array, the rowPerm argument of compute()/analyzePattern()
 * (:94,:392-400); NULL => identity.  It is stored and returned by rowsPermutation() only — as in
 * the reference, _solve_impl does not apply it (:266; callers do, test-qrkit.cpp:235,274). */
QRK_API int qrk_analyze_pattern(qrk_handle_t h, const int32_t* row_perm);
QRK_API int qrk_factorize(qrk_handle_t h);                                   /* on the blocks set above */
QRK_API int qrk_compute(qrk_handle_t h, const double* values, int memspace); /* set_blocks + analyze + factorize */
/* Fused factorize + Q^T b + back substitution + column permutation in ONE pass over the blocks
 * (what compute() followed by solve(b) returns; 16rc+8r+16c bytes per block instead of two passes). */
QRK_API int qrk_compute_solve(qrk_handle_t h, const double* values, const double* b, double* x, int memspace);
/* Same on blocks already resident (qrk_set_blocks / qrk_adopt_blocks); b, x in `memspace`. */
QRK_API int qrk_factorize_solve(qrk_handle_t h, const double* b, double* x, int memspace);

/* ---- accessors (BlockDiagonalSparseQR.h:108-112, 156-165, 242-254, 309-313) --------------------- */
QRK_API int qrk_rows(qrk_handle_t h, int64_t* rows);
QRK_API int qrk_cols(qrk_handle_t h, int64_t* cols);
QRK_API int qrk_rank(qrk_handle_t h, int64_t* rank);   /* = sum of block cols, as the reference (:439-444,:543) */
QRK_API int qrk_info(qrk_handle_t h, int32_t* info);   /* qrk_computation_info */
QRK_API int qrk_cols_permutation(qrk_handle_t h, int32_t* indices, int memspace);  /* int32[n_cols], P.indices() with A*P = Q*R (:519-521) */
QRK_API int qrk_rows_permutation(qrk_handle_t h, int32_t* indices, int memspace);  /* int32[n_rows] */

/* matrixR() (:156): column-major compressed sparse, n_rows x n_cols, int32 indices, exactly the
 * entries the reference emits (:475-479 FullQ, :496-500 BlockDiagonalQ) after setFromTriplets. */
QRK_API int qrk_matrix_r_nnz(qrk_handle_t h, int64_t* nnz);
QRK_API int qrk_matrix_r(qrk_handle_t h, int32_t* outer /* n_cols+1 */, int32_t* inner, double* values, int memspace);
/* matrixQ() (:235-237): the explicit row-major sparse Q, n_rows x n_rows, index rule :455-470
 * (FullQ) / :483-491 (BlockDiagonalQ) plus the identity tail (:530-533).  Built on demand from the
 * compact reflectors; the solve / apply paths never materialise it. */
QRK_API int qrk_matrix_q_nnz(qrk_handle_t h, int64_t* nnz);
QRK_API int qrk_matrix_q(qrk_handle_t h, int32_t* outer /* n_rows+1 */, int32_t* inner, double* values, int memspace);

/* Compact factors (the device-native form): packed V\R in the block-COO layout, tau[n_cols]. */
QRK_API int qrk_packed_factors(qrk_handle_t h, double* packed, double* tau, int memspace);

/* ---- matrixQ().transpose() * B and matrixQ() * B (:235-237 used at :266; BlockAngularSparseQR.h:365) */
/* B, Y: n_rows x nrhs column-major with leading dimensions ldb, ldy.  The result uses the index
 * layout of the reference's Q format (FullQ: thin part first, complement after n_cols, :455-470). */
QRK_API int qrk_apply_qt(qrk_handle_t h, const double* B, int64_t ldb, double* Y, int64_t ldy, int32_t nrhs, int memspace);
QRK_API int qrk_apply_q(qrk_handle_t h, const double* B, int64_t ldb, double* Y, int64_t ldy, int32_t nrhs, int memspace);
/* The thin factor Q1 = (A P) R^-1, n_rows x n_cols — the part of matrixQ() that _solve_impl (:266-267, topRows(rank)) and the
 * LM caller read: Y = Q1^T B (B: n_rows x nrhs, Y: n_cols x nrhs) and X = Q1 Y.  Defined for every solver kind (block
 * diagonal: FullQ layout only); for banded factors it is the only Q product there is (see below). */
QRK_API int qrk_apply_qt_thin(qrk_handle_t h, const double* B, int64_t ldb, double* Y, int64_t ldy, int32_t nrhs, int memspace);
QRK_API int qrk_apply_q_thin(qrk_handle_t h, const double* Y, int64_t ldy, double* X, int64_t ldx, int32_t nrhs, int memspace);

/* ---- solve (_solve_impl :258-280, solve :287-299) ------------------------------------------------ */
/* X = P * R^-1 * (Q^T B)[0:rank]; B: n_rows x nrhs (ldb), X: n_cols x nrhs (ldx). */
QRK_API int qrk_solve(qrk_handle_t h, const double* B, int64_t ldb, double* X, int64_t ldx, int32_t nrhs, int memspace);

/* ---- block angular: A = [J1 | J2] (BlockAngularSparseQR.h:79-281; BlockMatrix1x2.h:31-67) ----------------------
 * A handle of kind QRK_BLOCK_ANGULAR describes the LEFT block J1 (block diagonal, desc as above) and
 * border_cols = m2 columns of the dense RIGHT block J2.  The generic entry points then mean:
 *   qrk_compute / qrk_factorize      leftSolver.compute(J1); J2' = Q1^T J2; TSQR of the residual rows of J2'
 *                                    (replaces rightSolver.compute(Abot), :368, Right = ColPivHouseholderQR);
 *                                    R = [R1, Atop P2; 0, R2] (:285-308), P_c = [P1; m1 + P2] (:498-503)
 *   qrk_compute_solve / qrk_factorize_solve / qrk_solve   _solve_impl (:203-227); b: n, x: m1 + m2
 *   qrk_matrix_r(_nnz), qrk_cols_permutation, qrk_rank (= rank1 + rank2, :510), qrk_rows, qrk_cols (= m1 + m2)
 *   qrk_apply_qt / qrk_apply_q       matrixQ().transpose() * v = [I 0; 0 Q2^T] Q1^T v and matrixQ() * v = Q1 [I 0; 0 Q2] v on all
 *                                    n rows (_QProduct::evalTo, :598-644): the left factor in its FullQ layout, then the right
 *                                    solver's Q2 on the complement rows [m1, n).  On the fused TSQR path Q2 is built on demand
 *                                    from the residual panel kept by qrk_compute (not by the fused qrk_compute_solve:
 *                                    INVALID_ARGUMENT then), as the Householder QR of Abot P2 with its row signs aligned to R2
 * qrk_matrix_q / qrk_packed_factors keep referring to the LEFT factor Q1 (the reference's matrixQ() is an expression, not a matrix).
 * Uniform left blocks of 2x1, 3x1, 4x2, 7x2 with 1 <= m2 <= 8 and a ColPiv right solver take the fused in-SM TSQR path
 * (and support the multi-GPU exchange below); every other left block / border width / right solver takes the dense
 * right-block path (blocked compact-WY with DMMA, then ColPiv on the triangle), single GPU.
 * left_solver = QRK_LEFT_BANDED_BLOCKED: J1 is block banded (values = the slabs, as for kind QRK_BANDED_BLOCKED); Q1^T [J2 | b] is
 * the two-phase banded application INCLUDING its (extended) complement (the rows outside range(J1), which the dense right block
 * is factored on), x1 = R1^-1 (y1 - Atop x2) the banded back substitution.  qrk_apply_qt_thin gives [Q1thin^T b; z2] (the
 * vector _solve_impl back-substitutes); the n x n products qrk_apply_qt / qrk_apply_q return QRK_STATUS_UNSUPPORTED (banded
 * factors have no n x n Q here, see below).  Dense right-block path, single GPU. */
/* J2: n x m2 column-major with leading dimension ld (BlockMatrix1x2::rightBlock()).  Host: copied to the
 * device on the handle's stream; device: borrowed until the next compute returns. */
QRK_API int qrk_set_border(qrk_handle_t h, const double* J2, int64_t ld, int memspace);
/* Launch-bound block-angular steps (borders wider than 8 columns: ~250 small launches per compute; the three-launch TSQR step
 * of a narrow border on one GPU): from the second call with the same device buffers (values, rhs, x, border, stream) the
 * handle replays them from a CUDA graph it captured itself.  Set QRK_NO_GRAPH=1 to keep
 * every call eager; a stream that the CALLER is capturing is never captured again by the library. */
/* Multi-GPU (one handle per GPU, each owning a contiguous range of diagonal blocks and the matching rows of
 * J2 and b): with world_size > 1 the compute/solve calls stop after the per-GPU m2 x (m2+1) TSQR triangle;
 * the caller all-gathers the triangles (NCCL) and every rank calls qrk_angular_merge, which runs the TSQR
 * root redundantly and finishes the local part of the solution (x: m1_local + m2, the m2 shared
 * parameters last). */
QRK_API int qrk_angular_set_world(qrk_handle_t h, int32_t world_size);
QRK_API int qrk_angular_triangle_size(qrk_handle_t h, int64_t* doubles);
QRK_API int qrk_angular_local_triangle(qrk_handle_t h, double* tri, int memspace);
QRK_API int qrk_angular_merge(qrk_handle_t h, const double* tris, int32_t count, int memspace);
/* Fused exchange over NVLink peer memory (replaces the all-gather + qrk_angular_merge pair): after qrk_angular_p2p_attach the
 * compute / solve calls of a world > 1 handle are complete again — the TSQR root kernel stores this GPU's triangle into every
 * peer's exchange buffer, raises a flag, waits for the peers' flags (bounded spin) and merges the G triangles in rank order
 * (bit-identical shared parameters on every rank), all inside one launch.  Every rank must issue the same sequence of calls.
 *   qrk_angular_xchg_buffer   this handle's exchange buffer (a cudaMalloc allocation of *bytes bytes; call after set_world)
 *   qrk_ipc_export / _import  cudaIpcGetMemHandle / cudaIpcOpenMemHandle (64-byte handles) for one-process-per-GPU callers;
 *                             exchange the handles with any host-side collective (torch.distributed, MPI)
 *   qrk_angular_p2p_attach    peer_buffers[g] = rank g's exchange buffer as mapped in THIS process (own buffer at [rank])
 *                             — every rank must have returned from attach (host barrier) before any rank starts a step.
 *                             Attach also loads the kernels and allocates every buffer the later compute / solve calls of
 *                             this handle would allocate on first use (staging for host-memspace b / x / border / values, the
 *                             stored Abot panel): cudaMalloc waits for all running kernels of a device, so a first-use
 *                             allocation on one rank while another rank's root kernel on the SAME device waits for it
 *                             would stall both until the timeout (several handles per device: tests, ShardedBlockAngularSparseQR)
 *   qrk_angular_p2p_set_timeout  bound of the in-kernel wait for the peers' flags, in seconds (default 10); ranks must launch
 *                             their steps within this window of each other
 *   qrk_angular_p2p_status    *timed_out = 1 if a peer never arrived.  The condition is sticky until the next attach; the
 *                             merged triangle of that and every later step is poisoned with NaN (x, R2, the root record are
 *                             NaN, never a plausible wrong answer) and qrk_synchronize / qrk_rank / host-memspace solves
 *                             return QRK_STATUS_PEER_TIMEOUT */
QRK_API int qrk_angular_xchg_buffer(qrk_handle_t h, void** device_ptr, int64_t* bytes);
QRK_API int qrk_angular_p2p_attach(qrk_handle_t h, void* const* peer_buffers, int32_t world_size, int32_t rank);
QRK_API int qrk_angular_p2p_set_timeout(qrk_handle_t h, double seconds);
QRK_API int qrk_angular_p2p_status(qrk_handle_t h, int32_t* timed_out);
QRK_API int qrk_ipc_export(const void* device_ptr, void* handle64);
QRK_API int qrk_ipc_import(const void* handle64, void** device_ptr);
QRK_API int qrk_ipc_close(void* device_ptr);
/* One process driving several GPUs (no IPC needed: every cudaMalloc allocation is addressable once peer access is on):
 * cudaDeviceEnablePeerAccess(peer) issued on `device`; already enabled / device == peer is not an error.  Then pass the
 * handles' qrk_angular_xchg_buffer pointers straight to qrk_angular_p2p_attach (QRKit.hpp: ShardedBlockAngularSparseQR). */
QRK_API int qrk_enable_peer_access(int32_t device, int32_t peer);

/* ---- banded blocked (BandedBlockedSparseQR.h:122-344) ---------------------------------------------------------------
 * A handle of kind QRK_BANDED_BLOCKED describes num_blocks block rows of block_rows x block_cols; block row k sits at
 * rows [k*block_rows, ...) and columns [k*S, k*S + block_cols), S = block_cols - block_overlap
 * (fromBlockBandedPattern, SparseQRUtils.h:274-302).  values = the slabs, column-major, back to back.
 * Single GPU by construction (R of a fixed column order is a left-to-right recurrence, BandedBlockedSparseQR.h:463-508;
 * the device path reduces groups of block rows in parallel and keeps only a short chase sequential).  Generic entry points:
 *   qrk_compute / qrk_factorize / qrk_compute_solve / qrk_factorize_solve    factorize (:443-519) (+ fused solve)
 *   qrk_solve                         Q^T b by the window sweep (:655-670 / SparseBlockYTY.h:102-139), banded back
 *                                     substitution (:299-304)
 *   qrk_apply_qt_thin / _q_thin       Q1^T b (n_cols values) and Q1 y (n_rows values), Q1 = A R^-1 the thin factor (unique up
 *                                     to column signs): what solve and the LM caller use of matrixQ()
 *   qrk_apply_qt / qrk_apply_q        QRK_STATUS_UNSUPPORTED: the two-phase factorisation represents Q as an isometry into an
 *                                     EXTENDED complement (every group's overlap rows enter as virtual zero rows), so there
 *                                     is no n x n orthogonal matrix to multiply with (the reference's complement depends on
 *                                     its own window blocking and is not unique either)
 *   qrk_matrix_r(_nnz)                R as CSC in the reference's exact stored pattern: per merged window (the blocks its block
 *                                     detection + mergeBlocks yield for this slab geometry, SparseQRUtils.h:186-253, 308-385) the
 *                                     dense rectangle of solved rows x window columns, explicit zeros included (:484-491);
 *                                     values as the reference up to row signs
 *   qrk_rank (= cols, :514), qrk_cols_permutation (identity), qrk_rows_permutation
 * Slab shapes (block_rows, block_cols, overlap) with the fast two-phase kernels: (16,24,16), (7,4,2), (7,2,0), (8,8,4), (12,8,4),
 * (4,6,4); every other shape runs on the general window chain below (same results, sequential, with an exact n x n Q). */

/* BandedBlockedSparseQR on a GENERAL banded matrix (the analyzePattern else-branch, BandedBlockedSparseQR.h:408-426: row ordering +
 * block detection): the caller orders the rows (qrk_order_as_banded_as_possible), detects the blocks (qrk_detect_band_starts, or
 * qrk_detect_blocks for the reference's merged windows), extracts them (qrk_extract_blocks: dense, column-major, back to back =
 * the `values` of qrk_compute / qrk_compute_solve) and creates the handle from the block list {idxRow, idxCol, numRows, numCols}:
 * rows contiguous and in order, first columns non-decreasing.  Any block sizes, overlaps and column steps; windows run
 * sequentially on one SM (banded_generic.cuh), so this path is general, not fast.  qrk_apply_qt / qrk_apply_q are an exact
 * n x n Q here: [thin part (n_cols) ; complement (n_rows - n_cols, window order)].  matrixR(): the reference's stored pattern
 * for the windows mergeBlocks makes of this block list (suggested_block_cols: 0 = 2).  The same general chain (one window per
 * slab) serves qrk_create for slab shapes outside the instantiated list. */
QRK_API int qrk_create_banded_general(const int32_t* blocks, int64_t num_blocks, int64_t n_rows, int64_t n_cols, int32_t device,
                                      int32_t suggested_block_cols, qrk_handle_t* out);

/* ---- pattern analysis on the host (no GPU needed; qrkit_b200/csrc/structure.cpp) ---------------------- */
/* SparseQROrdering::AsBandedAsPossible (SparseQROrdering.h:53-120): rows of a row-major (CSR) pattern stably sorted by
 * their first stored column.  perm_indices[original row] = new row (PermutationMatrix::indices()); *has_permutation = 0
 * when the rows already are in order (then perm_indices is the identity). */
QRK_API int qrk_order_as_banded_as_possible(int64_t rows, int64_t cols, const int32_t* csr_outer, const int32_t* csr_inner,
                                            int32_t* perm_indices, int32_t* has_permutation);
/* SparseQROrdering::ColumnDensity (SparseQROrdering.h:22-50): columns of a column-major (CSC) pattern stably sorted by
 * their number of stored entries, ascending; perm_indices[original column] = new column. */
QRK_API int qrk_order_column_density(int64_t cols, const int32_t* csc_outer, int32_t* perm_indices);
/* BlockBandedMatrixInfo::operator() (SparseQRUtils.h:186-253) including mergeBlocks (:308-385), on a row-major pattern
 * whose rows are already ordered.  blocks: 4 int32 per block {idxRow, idxCol, numRows, numCols} in blockOrder order;
 * blocks == NULL only counts.  suggested_block_cols = the SuggestedBlockCols template argument (default 2). */
QRK_API int qrk_detect_blocks(int64_t rows, int64_t cols, const int32_t* csr_outer, const int32_t* csr_inner,
                              int32_t suggested_block_cols, int32_t* blocks, int64_t capacity, int64_t* num_blocks,
                              int64_t* nonzero_q_estimate);
/* The same detection WITHOUT mergeBlocks: one block per distinct band start (first stored column), rows in order — the
 * window chain qrk_create_banded_general factors (its own windows need not be portrait: rows still to be finalised are
 * carried from window to window). */
QRK_API int qrk_detect_band_starts(int64_t rows, int64_t cols, const int32_t* csr_outer, const int32_t* csr_inner, int32_t* blocks,
                                   int64_t capacity, int64_t* num_blocks);
/* fromBlockDiagonalPattern (SparseQRUtils.h:255-272) and fromBlockBandedPattern (:274-302, merge included). */
QRK_API int qrk_block_diagonal_pattern(int64_t rows, int64_t cols, int32_t block_rows, int32_t block_cols, int32_t* blocks,
                                       int64_t capacity, int64_t* num_blocks);
QRK_API int qrk_block_banded_pattern(int64_t rows, int64_t cols, int32_t block_rows, int32_t block_cols, int32_t block_overlap,
                                     int32_t suggested_block_cols, int32_t* blocks, int64_t capacity, int64_t* num_blocks);
/* The mat.block(idxRow, idxCol, numRows, numCols) loop of SparseBlockDiagonal::fromSparseMatrix (SparseBlockDiagonal.h:
 * 123-128): dense blocks of P*A in the block-COO layout (column-major, back to back), P given as row_perm[original row]
 * = new row (NULL = identity).  Note: the reference slices the UNPERMUTED matrix there (:127) although the blocks were
 * detected on the permuted one; pass row_perm = NULL to reproduce that. */
QRK_API int qrk_extract_blocks(int64_t rows, int64_t cols, const int32_t* csc_outer, const int32_t* csc_inner,
                               const double* csc_values, const int32_t* row_perm, const int32_t* blocks, int64_t num_blocks,
                               double* values_out);

/* ---- measurement hooks ---------------------------------------------------------------------------- */
/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
QRK_API int qrk_launch_count(qrk_handle_t h, int64_t* launches);
/* Fill nb uniform r x c blocks (or a vector when c == 0) on the device with the shared counter-based
 * generator of the test-suite (splitmix64, U[lo,hi)); lets the bench create HBM-resident inputs
 * without a host copy. */
QRK_API int qrk_synth_fill(double* device_out, uint64_t seed, int64_t block0, int64_t nb, int32_t r, int32_t c,
                           double lo, double hi, void* cuda_stream);

/* ---- device-side assembly for the LM caller ------------------------------------------------
 * The reference's published benchmark is a whole Levenberg-Marquardt ellipse fit (README.md:25-30); its functor builds the
 * Jacobian on the host every iteration (examples/ellipse_fitting.cpp:85-113 = bench/bench_sparse_qr_extra.cpp:79-114) and
 * the solver is constructed from it (:126-141).  With the factorisation at ~60 us, a host-built Jacobian (248 MB at 1M
 * points over PCIe) would dominate; these two entry points keep the whole iteration on the device: the block-COO left
 * block J1 (n blocks 2x1), the dense border J2 (2n x 5, column-major, ld = 2n) and the right-hand side are written straight
 * into the buffers qrk_set_border / qrk_compute_solve(QRK_DEVICE) take.  All pointers are device pointers.
 *   params: t[0..n) followed by {a, b, x0, y0, r}  (the functor's input vector, bench :68-76)
 *   rhs   : -f(params)  (so that the least-squares solution is the Gauss-Newton step to ADD to params)
 *   cost  : *cost += sum f^2  (optional; zero it first) */
QRK_API int qrk_ellipse_points(double* px, double* py, int64_t n, double a, double b, double x0, double y0, double r,
                               void* cuda_stream);   /* synthetic samples over 1.3 pi of the ellipse (bench :277-282) */
QRK_API int qrk_ellipse_assemble(const double* px, const double* py, const double* params, int64_t n, double* J1, double* J2,
                                 double* rhs, double* cost, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* QRKIT_B200_H_ */
