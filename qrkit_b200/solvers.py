"""Python mirror of the reference's solver interface over the C ABI (test / bench harness).

The product host façade is C++ (include/qrkit_b200/QRKit.hpp); this module gives the parity tests
the same vocabulary as the reference's own tests (test/test-qrkit.cpp:167-206): ``compute``,
``matrixQ``, ``matrixR``, ``solve``, ``rank``, ``info``, ``colsPermutation``, ``rowsPermutation``.
Every call goes through libqrkit_b200.so; nothing here computes."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import (QRK_BLOCK_DIAGONAL, QRK_DEVICE, QRK_FULL_Q, QRK_HOST, QRK_PIVOT_COLPIV, QRK_PIVOT_NONE, QrkDesc,
                   check, lib)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))      # raw device pointer (e.g. torch.Tensor.data_ptr())


class SparseBlockDiagonal:
    """Input container (SparseBlockDiagonal.h:44-163): the block-COO arrays.

    ``values``: flat float64, block i column-major at offset sum_{l<i} r_l*c_l.
    Uniform blocks: pass ``block_rows``/``block_cols`` (fromBlockDiagonalPattern, :72-89)."""

    def __init__(self, values, *, num_blocks=None, block_rows=0, block_cols=0, rows=None, cols=None, n_rows=0, n_cols=0):
        self.values = values if not isinstance(values, np.ndarray) else np.ascontiguousarray(values, dtype=np.float64)
        if rows is not None:
            self.block_sizes_rows = np.ascontiguousarray(rows, dtype=np.int32)
            self.block_sizes_cols = np.ascontiguousarray(cols, dtype=np.int32)
            self.num_blocks = len(self.block_sizes_rows)
            self.block_rows = self.block_cols = 0
            self._rows = int(self.block_sizes_rows.sum())
            self._cols = int(self.block_sizes_cols.sum())
        else:
            self.block_sizes_rows = self.block_sizes_cols = None
            self.block_rows, self.block_cols = int(block_rows), int(block_cols)
            if num_blocks is None:
                num_blocks = len(values) // (block_rows * block_cols)
            self.num_blocks = int(num_blocks)
            self._rows = self.num_blocks * self.block_rows
            self._cols = self.num_blocks * self.block_cols
        self.n_rows = int(n_rows) or self._rows
        self.n_cols = int(n_cols) or self._cols

    def rows(self):
        return self.n_rows

    def cols(self):
        return self.n_cols

    def size(self):
        return self.num_blocks


class SparseOut:
    """Compressed sparse matrix as the C ABI returns it (int32 outer/inner, float64 values)."""

    def __init__(self, rows, cols, outer, inner, values, row_major):
        self.rows, self.cols, self.outer, self.inner, self.values, self.row_major = rows, cols, outer, inner, values, row_major

    def tocsc(self):
        import scipy.sparse as sp
        cls = sp.csr_matrix if self.row_major else sp.csc_matrix
        return cls((self.values, self.inner, self.outer), shape=(self.rows, self.cols))

    def toarray(self):
        return self.tocsc().toarray()


class BlockDiagonalSparseQR:
    """BlockDiagonalSparseQR<BlockQRSolver, QFormat> (reference BlockDiagonalSparseQR.h:37-335).

    ``pivoting`` selects the per-block dense solver: QRK_PIVOT_COLPIV = ColPivHouseholderQR (the
    reference's tests, test-qrkit.cpp:49-51), QRK_PIVOT_NONE = HouseholderQR."""

    def __init__(self, mat: SparseBlockDiagonal | None = None, *, pivoting=QRK_PIVOT_COLPIV, q_format=QRK_FULL_Q, device=0,
                 stream=None):
        self._h = C.c_void_p()
        self._pivoting, self._q_format, self._device, self._stream = pivoting, q_format, device, stream
        self._shape_key = None
        self._mat = None
        if mat is not None:
            self.compute(mat)

    # ---- life cycle -------------------------------------------------------------------------
    def _ensure_handle(self, mat: SparseBlockDiagonal):
        key = (mat.num_blocks, mat.block_rows, mat.block_cols, mat.n_rows, mat.n_cols,
               None if mat.block_sizes_rows is None else (mat.block_sizes_rows.tobytes(), mat.block_sizes_cols.tobytes()))
        if self._h and key == self._shape_key:
            return
        self.close()
        d = QrkDesc()
        d.kind, d.device, d.num_blocks = QRK_BLOCK_DIAGONAL, self._device, mat.num_blocks
        d.block_rows, d.block_cols = mat.block_rows, mat.block_cols
        if mat.block_sizes_rows is not None:
            d.rows = mat.block_sizes_rows.ctypes.data_as(C.POINTER(C.c_int32))
            d.cols = mat.block_sizes_cols.ctypes.data_as(C.POINTER(C.c_int32))
        d.n_rows, d.n_cols = mat.n_rows, mat.n_cols
        d.pivoting, d.q_format = self._pivoting, self._q_format
        h = C.c_void_p()
        check(lib().qrk_create(C.byref(d), C.byref(h)))
        self._h = h
        self._shape_key = key
        if self._stream is not None:
            check(lib().qrk_set_stream(self._h, C.c_void_p(int(self._stream))), self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib().qrk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _space(a):
        return QRK_HOST if isinstance(a, np.ndarray) else QRK_DEVICE

    # ---- compute (:94-104) -------------------------------------------------------------------
    def compute(self, mat: SparseBlockDiagonal, row_perm=None):
        self._ensure_handle(mat)
        self._mat = mat
        self.analyzePattern(mat, row_perm)
        self.factorize(mat)
        return self

    def analyzePattern(self, mat: SparseBlockDiagonal, row_perm=None):
        self._ensure_handle(mat)
        self._mat = mat
        rp = None if row_perm is None else np.ascontiguousarray(row_perm, dtype=np.int32)
        check(lib().qrk_analyze_pattern(self._h, _ptr(rp)), self._h)

    def factorize(self, mat: SparseBlockDiagonal):
        self._ensure_handle(mat)
        self._mat = mat
        check(lib().qrk_set_blocks(self._h, _ptr(mat.values), self._space(mat.values)), self._h)
        check(lib().qrk_factorize(self._h), self._h)

    def compute_solve(self, mat: SparseBlockDiagonal, b):
        """Fused compute(mat) + solve(b) in one pass over the blocks."""
        self._ensure_handle(mat)
        self._mat = mat
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(mat.n_cols)
        check(lib().qrk_compute_solve(self._h, _ptr(mat.values), _ptr(b), _ptr(x), QRK_HOST), self._h)
        return x

    # ---- accessors ----------------------------------------------------------------------------
    def rows(self):
        v = C.c_int64()
        check(lib().qrk_rows(self._h, C.byref(v)), self._h)
        return v.value

    def cols(self):
        v = C.c_int64()
        check(lib().qrk_cols(self._h, C.byref(v)), self._h)
        return v.value

    def rank(self):
        v = C.c_int64()
        check(lib().qrk_rank(self._h, C.byref(v)), self._h)
        return v.value

    def info(self):
        v = C.c_int32()
        check(lib().qrk_info(self._h, C.byref(v)), self._h)
        return v.value

    def colsPermutation(self):
        p = np.empty(self.cols(), dtype=np.int32)
        check(lib().qrk_cols_permutation(self._h, _ptr(p), QRK_HOST), self._h)
        return p

    def rowsPermutation(self):
        p = np.empty(self.rows(), dtype=np.int32)
        check(lib().qrk_rows_permutation(self._h, _ptr(p), QRK_HOST), self._h)
        return p

    def matrixR(self):
        nnz = C.c_int64()
        check(lib().qrk_matrix_r_nnz(self._h, C.byref(nnz)), self._h)
        outer = np.empty(self.cols() + 1, dtype=np.int32)
        inner = np.empty(nnz.value, dtype=np.int32)
        vals = np.empty(nnz.value)
        check(lib().qrk_matrix_r(self._h, _ptr(outer), _ptr(inner), _ptr(vals), QRK_HOST), self._h)
        return SparseOut(self.rows(), self.cols(), outer, inner, vals, False)

    def matrixQ(self):
        nnz = C.c_int64()
        check(lib().qrk_matrix_q_nnz(self._h, C.byref(nnz)), self._h)
        outer = np.empty(self.rows() + 1, dtype=np.int32)
        inner = np.empty(nnz.value, dtype=np.int32)
        vals = np.empty(nnz.value)
        check(lib().qrk_matrix_q(self._h, _ptr(outer), _ptr(inner), _ptr(vals), QRK_HOST), self._h)
        return SparseOut(self.rows(), self.rows(), outer, inner, vals, True)

    def packed(self):
        n = C.c_int64()
        check(lib().qrk_total_values(self._h, C.byref(n)), self._h)
        pk = np.empty(n.value)
        tau = np.empty(self.cols())
        check(lib().qrk_packed_factors(self._h, _ptr(pk), _ptr(tau), QRK_HOST), self._h)
        return pk, tau

    # ---- matrixQ().transpose() * B, matrixQ() * B, solve(B) -----------------------------------------
    def _apply(self, fn, B, out_rows):
        B = np.asarray(B, dtype=np.float64)
        vec = B.ndim == 1
        Bf = np.asfortranarray(B.reshape(len(B), 1) if vec else B)
        nrhs = Bf.shape[1]
        Y = np.empty((out_rows, nrhs), order="F")
        check(fn(self._h, _ptr(Bf), Bf.shape[0], _ptr(Y), out_rows, nrhs, QRK_HOST), self._h)
        return Y[:, 0].copy() if vec else Y

    def applyQt(self, B):
        """matrixQ().transpose() * B"""
        return self._apply(lib().qrk_apply_qt, B, self.rows())

    def applyQ(self, B):
        """matrixQ() * B"""
        return self._apply(lib().qrk_apply_q, B, self.rows())

    def applyQtThin(self, B):
        """(matrixQ().transpose() * B).topRows(cols()): Q1^T B with the thin factor Q1 = (A P) R^-1 (qrk_apply_qt_thin)"""
        return self._apply(lib().qrk_apply_qt_thin, B, self.cols())

    def applyQThin(self, Y):
        """matrixQ() * [Y; 0]: Q1 Y (qrk_apply_q_thin)"""
        return self._apply(lib().qrk_apply_q_thin, Y, self.rows())

    def solve(self, B):
        return self._apply(lib().qrk_solve, B, self.cols())

    def launch_count(self):
        v = C.c_int64()
        check(lib().qrk_launch_count(self._h, C.byref(v)), self._h)
        return v.value


class BandedSlabs:
    """A block-banded left block as the device path takes it: num_blocks dense slabs of block_rows x block_cols (column-major,
    back to back), slab k at rows [k*block_rows, ...), columns [k*(block_cols-overlap), ...) — the matrix of
    fromBlockBandedPattern (SparseQRUtils.h:274-302).  n_cols: columns of the matrix when it ends inside the last slab."""

    def __init__(self, values, *, num_blocks, block_rows, block_cols, overlap, n_cols=0):
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        self.num_blocks, self.block_rows, self.block_cols, self.overlap = num_blocks, block_rows, block_cols, overlap
        self.n_cols = n_cols or (num_blocks - 1) * (block_cols - overlap) + block_cols

    def rows(self):
        return self.num_blocks * self.block_rows

    def cols(self):
        return self.n_cols


class BlockMatrix1x2:
    """BlockMatrix1x2<Left, Right> (BlockMatrix1x2.h:31-67): references to a left block-diagonal matrix and a
    dense right block (n x m2, column-major) with the same number of rows."""

    def __init__(self, left: SparseBlockDiagonal, right):
        self.left = left
        self.right = np.asfortranarray(right, dtype=np.float64) if isinstance(right, np.ndarray) else right
        if isinstance(right, np.ndarray):
            assert self.right.shape[0] == left.rows(), "blocks must have the same number of rows (BlockMatrix1x2.h:37)"

    def leftBlock(self):
        return self.left

    def rightBlock(self):
        return self.right

    def rows(self):
        return self.left.rows()

    def cols(self):
        return self.left.cols() + self.right.shape[1]


class BlockAngularSparseQR(BlockDiagonalSparseQR):
    """BlockAngularSparseQR<LeftSolver, RightSolver> (BlockAngularSparseQR.h:79-281).  mat.left is a SparseBlockDiagonal
    (LeftSolver = BlockDiagonalSparseQR: narrow borders go through the fused TSQR kernels, Eigen's ColPiv rule at the root) or a
    BandedSlabs (LeftSolver = BandedBlockedSparseQR, the pair of test/test-qrkit.cpp:44-48); right_solver: 0 =
    ColPivHouseholderQR<MatrixXd>, 1 = BlockedThinDenseQR / HouseholderQR (unpivoted), 2 = BlockedThinSparseQR (ColPiv inside
    panels of `right_panel` columns, zero-pivot columns deferred; test/test-qrkit.cpp:54-57)."""

    def __init__(self, mat: BlockMatrix1x2 | None = None, *, pivoting=QRK_PIVOT_COLPIV, device=0, stream=None, world=1,
                 right_solver=0, right_panel=0):
        self._world = world
        self._right_solver = right_solver      # 0: ColPivHouseholderQR<MatrixXd>, 1: BlockedThinDenseQR / HouseholderQR (unpivoted), 2: BlockedThinSparseQR
        self._right_panel = right_panel        # SuggestedBlockCols of BlockedThinSparseQR (0 = the reference's 2)
        super().__init__(None, pivoting=pivoting, q_format=QRK_FULL_Q, device=device, stream=stream)
        if mat is not None:
            self.compute(mat)

    def _ensure_handle_angular(self, mat: BlockMatrix1x2):
        left, m2 = mat.left, mat.right.shape[1]
        banded = isinstance(left, BandedSlabs)      # LeftSolver = BandedBlockedSparseQR (test/test-qrkit.cpp:44-48)
        key = ("angular", left.num_blocks, left.block_rows, left.block_cols, m2, self._right_solver, self._right_panel,
               (left.overlap, left.n_cols) if banded else None)
        if self._h and key == self._shape_key:
            return
        self.close()
        d = QrkDesc()
        d.kind, d.device, d.num_blocks = capi.QRK_BLOCK_ANGULAR, self._device, left.num_blocks
        d.block_rows, d.block_cols = left.block_rows, left.block_cols
        d.pivoting, d.q_format, d.border_cols = self._pivoting, QRK_FULL_Q, m2
        d.right_solver = self._right_solver
        d.reserved[1] = self._right_panel
        if banded:
            d.left_solver, d.block_overlap, d.n_cols, d.pivoting = 1, left.overlap, left.n_cols + m2, QRK_PIVOT_NONE
        h = C.c_void_p()
        check(lib().qrk_create(C.byref(d), C.byref(h)))
        self._h, self._shape_key = h, key
        if self._stream is not None:
            check(lib().qrk_set_stream(self._h, C.c_void_p(int(self._stream))), self._h)
        if self._world > 1:
            check(lib().qrk_angular_set_world(self._h, self._world), self._h)

    def compute(self, mat: BlockMatrix1x2):
        self._ensure_handle_angular(mat)
        self._mat = mat
        check(lib().qrk_set_border(self._h, _ptr(mat.right), mat.right.shape[0], QRK_HOST), self._h)
        check(lib().qrk_compute(self._h, _ptr(mat.left.values), QRK_HOST), self._h)
        return self

    def compute_solve(self, mat: BlockMatrix1x2, b):
        self._ensure_handle_angular(mat)
        self._mat = mat
        b = np.ascontiguousarray(b, dtype=np.float64)
        self._x = np.empty(mat.cols())
        check(lib().qrk_set_border(self._h, _ptr(mat.right), mat.right.shape[0], QRK_HOST), self._h)
        check(lib().qrk_compute_solve(self._h, _ptr(mat.left.values), _ptr(b), _ptr(self._x), QRK_HOST), self._h)
        return self._x          # world > 1: filled by merge()

    def solve(self, B):
        if self._world > 1:
            B = np.ascontiguousarray(B, dtype=np.float64)
            self._x = np.empty(self.cols())
            check(lib().qrk_solve(self._h, _ptr(B), len(B), _ptr(self._x), self.cols(), 1, QRK_HOST), self._h)
            return self._x
        return super().solve(B)

    # ---- multi-GPU TSQR exchange ------------------------------------------------------------------
    def triangle_size(self):
        v = C.c_int64()
        check(lib().qrk_angular_triangle_size(self._h, C.byref(v)), self._h)
        return v.value

    def local_triangle(self):
        t = np.empty(self.triangle_size())
        check(lib().qrk_angular_local_triangle(self._h, _ptr(t), QRK_HOST), self._h)
        return t

    def merge(self, triangles):
        t = np.ascontiguousarray(triangles, dtype=np.float64).reshape(-1)
        check(lib().qrk_angular_merge(self._h, _ptr(t), len(t) // self.triangle_size(), QRK_HOST), self._h)
        return getattr(self, "_x", None)


class BandedBlockedSparseQR(BlockDiagonalSparseQR):
    """BandedBlockedSparseQR<SparseMatrix, HouseholderQR<MatrixXd>, BlockOverlap, ...> (BandedBlockedSparseQR.h:122-344) for
    the fixed block-banded pattern: ``slabs`` is the block-COO array of nb dense block_rows x block_cols slabs (column-major),
    slab k at rows [k*block_rows, ...), columns [k*(block_cols-overlap), ...)."""

    def __init__(self, slabs=None, *, num_blocks=None, block_rows=0, block_cols=0, overlap=0, n_cols=0, device=0, stream=None):
        self._geom = (block_rows, block_cols, overlap)
        self._n_cols = n_cols                # 0: full last slab; else the matrix ends inside the last slab
        super().__init__(None, pivoting=QRK_PIVOT_NONE, q_format=QRK_FULL_Q, device=device, stream=stream)
        if slabs is not None:
            self.compute(slabs, num_blocks)

    def _ensure_handle_banded(self, nb):
        br, bc, ov = self._geom
        key = ("banded", nb, br, bc, ov)
        if self._h and key == self._shape_key:
            return
        self.close()
        d = QrkDesc()
        d.kind, d.device, d.num_blocks = capi.QRK_BANDED_BLOCKED, self._device, nb
        d.block_rows, d.block_cols, d.block_overlap = br, bc, ov
        d.n_cols = self._n_cols
        h = C.c_void_p()
        check(lib().qrk_create(C.byref(d), C.byref(h)))
        self._h, self._shape_key = h, key
        if self._stream is not None:
            check(lib().qrk_set_stream(self._h, C.c_void_p(int(self._stream))), self._h)

    @classmethod
    def from_sparse(cls, A, *, merged=False, suggested_block_cols=2, device=0, stream=None):
        """compute(const SparseMatrix&) on a GENERAL banded matrix — the analyzePattern else-branch of the reference
        (BandedBlockedSparseQR.h:408-426): AsBandedAsPossible row ordering, block detection, dense extraction of the blocks, then the
        general window chain (qrk_create_banded_general).  merged=True factors the reference's merged windows instead of one
        window per band start (same R; bigger windows).  rowsPermutation() returns the ordering; as in the reference, solve() expects
        the right-hand side already permuted (test/test-qrkit.cpp:235)."""
        from . import structure
        perm, has = structure.as_banded_as_possible(A)
        Ar = A.tocsr()
        if has:
            inv = np.empty_like(perm)
            inv[perm] = np.arange(len(perm), dtype=np.int32)
            Ar = Ar[inv, :]
        blocks = structure.detect_blocks(Ar, suggested_block_cols)[0] if merged else structure.detect_band_starts(Ar)
        values = structure.extract_blocks(A, blocks, perm if has else None)
        self = cls(block_rows=0, block_cols=0, overlap=0, device=device, stream=stream)
        blocks = np.ascontiguousarray(blocks, dtype=np.int32)
        h = C.c_void_p()
        check(lib().qrk_create_banded_general(_ptr(blocks), len(blocks), A.shape[0], A.shape[1], device, suggested_block_cols, C.byref(h)))
        self._h, self._shape_key = h, ("banded-general", len(blocks))
        if stream is not None:
            check(lib().qrk_set_stream(self._h, C.c_void_p(int(stream))), self._h)
        check(lib().qrk_compute(self._h, _ptr(values), QRK_HOST), self._h)
        check(lib().qrk_analyze_pattern(self._h, _ptr(np.ascontiguousarray(perm, dtype=np.int32))), self._h)   # rowsPermutation()
        self.blocks, self.values = blocks, values
        return self

    def _nb(self, slabs, num_blocks):
        br, bc, _ = self._geom
        return int(num_blocks) if num_blocks is not None else len(slabs) // (br * bc)

    def compute(self, slabs, num_blocks=None):
        slabs = np.ascontiguousarray(slabs, dtype=np.float64)
        self._ensure_handle_banded(self._nb(slabs, num_blocks))
        check(lib().qrk_compute(self._h, _ptr(slabs), QRK_HOST), self._h)
        return self

    def compute_solve(self, slabs, b, num_blocks=None):
        slabs = np.ascontiguousarray(slabs, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        self._ensure_handle_banded(self._nb(slabs, num_blocks))
        x = np.empty(self.cols())
        check(lib().qrk_compute_solve(self._h, _ptr(slabs), _ptr(b), _ptr(x), QRK_HOST), self._h)
        return x
