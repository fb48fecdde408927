"""ctypes binding of the host-side pattern analysis (include/qrkit_b200.h, "pattern analysis on the host"): row ordering,
block detection and block extraction in front of the hot path.  Computes nothing itself; mirrors
SparseQROrdering / SparseQRUtils::BlockBandedMatrixInfo / SparseBlockDiagonal::fromSparseMatrix of the reference."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import check, lib


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def _csr(A):
    A = A.tocsr()
    A.sort_indices()
    return A, np.ascontiguousarray(A.indptr, dtype=np.int32), np.ascontiguousarray(A.indices, dtype=np.int32)


def as_banded_as_possible(A):
    """AsBandedAsPossible (SparseQROrdering.h:53-120).  Returns (perm_indices, has_permutation); perm_indices[orig] = new row."""
    A, outer, inner = _csr(A)
    perm = np.empty(A.shape[0], dtype=np.int32)
    has = C.c_int32(0)
    check(lib().qrk_order_as_banded_as_possible(A.shape[0], A.shape[1], _p(outer), _p(inner), _p(perm), C.byref(has)))
    return perm, bool(has.value)


def column_density(A):
    """ColumnDensity (SparseQROrdering.h:22-50): perm_indices[orig column] = new column (ascending nnz, stable)."""
    A = A.tocsc()
    outer = np.ascontiguousarray(A.indptr, dtype=np.int32)
    perm = np.empty(A.shape[1], dtype=np.int32)
    check(lib().qrk_order_column_density(A.shape[1], _p(outer), _p(perm)))
    return perm


def _blocks_call(fn, *args):
    n = C.c_int64(0)
    check(fn(*args, None, 0, C.byref(n)))
    out = np.zeros((max(n.value, 1), 4), dtype=np.int32)
    check(fn(*args, _p(out), n.value, C.byref(n)))
    return out[:n.value]


def detect_blocks(A, suggested_block_cols=2):
    """BlockBandedMatrixInfo::operator() + mergeBlocks (SparseQRUtils.h:186-253, 308-385) on a row-ordered matrix.
    Returns (blocks[nb, 4] = idxRow, idxCol, numRows, numCols; nonZeroQEstimate)."""
    A, outer, inner = _csr(A)
    n = C.c_int64(0)
    nzq = C.c_int64(0)
    L = lib()
    check(L.qrk_detect_blocks(A.shape[0], A.shape[1], _p(outer), _p(inner), suggested_block_cols, None, 0, C.byref(n), C.byref(nzq)))
    out = np.zeros((max(n.value, 1), 4), dtype=np.int32)
    check(L.qrk_detect_blocks(A.shape[0], A.shape[1], _p(outer), _p(inner), suggested_block_cols, _p(out), n.value, C.byref(n), C.byref(nzq)))
    return out[:n.value], nzq.value


def detect_band_starts(A):
    """The same detection without mergeBlocks: one block per distinct band start (qrk_detect_band_starts)."""
    A, outer, inner = _csr(A)
    return _blocks_call(lib().qrk_detect_band_starts, A.shape[0], A.shape[1], _p(outer), _p(inner))


def block_diagonal_pattern(rows, cols, block_rows, block_cols):
    return _blocks_call(lib().qrk_block_diagonal_pattern, rows, cols, block_rows, block_cols)


def block_banded_pattern(rows, cols, block_rows, block_cols, overlap, suggested_block_cols=2):
    return _blocks_call(lib().qrk_block_banded_pattern, rows, cols, block_rows, block_cols, overlap, suggested_block_cols)


def extract_blocks(A, blocks, row_perm=None):
    """Dense blocks of P*A in the block-COO layout (SparseBlockDiagonal.h:123-128)."""
    A = A.tocsc()
    A.sort_indices()
    outer = np.ascontiguousarray(A.indptr, dtype=np.int32)
    inner = np.ascontiguousarray(A.indices, dtype=np.int32)
    vals = np.ascontiguousarray(A.data, dtype=np.float64)
    blocks = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 4)
    out = np.empty(int((blocks[:, 2].astype(np.int64) * blocks[:, 3]).sum()), dtype=np.float64)
    rp = None if row_perm is None else np.ascontiguousarray(row_perm, dtype=np.int32)
    check(lib().qrk_extract_blocks(A.shape[0], A.shape[1], _p(outer), _p(inner), _p(vals), _p(rp), _p(blocks), len(blocks), _p(out)))
    return out


def from_sparse_matrix(A, suggested_block_cols=2):
    """SparseBlockDiagonal::fromSparseMatrix (SparseBlockDiagonal.h:96-130): row ordering, block detection, block
    extraction from the row-ordered matrix.  Returns (values, block_rows[], block_cols[], blocks, row_perm, has_permutation)."""
    perm, has = as_banded_as_possible(A)
    Ar = A.tocsr()
    if has:
        inv = np.empty_like(perm)
        inv[perm] = np.arange(len(perm), dtype=np.int32)
        Ar = Ar[inv, :]                                  # rows of P*A: new row i is original row inv[i]
    blocks, _ = detect_blocks(Ar, suggested_block_cols)
    vals = extract_blocks(A, blocks, perm if has else None)
    return vals, blocks[:, 2].copy(), blocks[:, 3].copy(), blocks, perm, has
