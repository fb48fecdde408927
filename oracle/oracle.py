"""ctypes front-end of the CPU ORACLE (oracle/qrkit_oracle.hpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py.  The product package (qrkit_b200/) never imports this.

PARITY UNPINNED: the reference cannot be compiled here (Eigen absent) and its tests hold no golden
vectors; see the header of qrkit_oracle.hpp for how the oracle is cross-checked instead.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libqrkit_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (seconds).  Returns the path of the shared library."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(os.path.join(_HERE, f)) for f in ("qrkit_oracle.hpp", "oracle_capi.cpp")
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_bd_factorize.restype = C.c_void_p
        L.orc_banded_factorize.restype = C.c_void_p
        L.orc_angular_factorize_bd.restype = C.c_void_p
        L.orc_angular_factorize_banded.restype = C.c_void_p
        for name in ("orc_bd_q_nnz", "orc_bd_r_nnz", "orc_banded_r_nnz", "orc_angular_r_nnz"):
            getattr(L, name).restype = C.c_int64
        for name in ("orc_bd_compact_uniform", "orc_bd_reference_uniform", "orc_angular_reference_uniform",
                     "orc_synth_value"):
            getattr(L, name).restype = C.c_double
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ------------------------------------------------------------------------------ dense kernels
def householder_qr(A):
    """Eigen::HouseholderQR restatement.  A: (r, c) array.  Returns (packed (r,c), tau)."""
    A = np.asarray(A, dtype=np.float64)
    r, c = A.shape
    buf = np.asfortranarray(A.copy())
    tau = np.zeros(min(r, c))
    lib().orc_householder_qr(_d(buf), r, c, _d(tau))
    return buf, tau


def colpiv_qr(A):
    """Eigen::ColPivHouseholderQR restatement.  Returns (packed, tau, perm, nonzero_pivots)."""
    A = np.asarray(A, dtype=np.float64)
    r, c = A.shape
    buf = np.asfortranarray(A.copy())
    tau = np.zeros(min(r, c))
    perm = np.zeros(c, dtype=np.int32)
    nz = lib().orc_colpiv_qr(_d(buf), r, c, _d(tau), _i(perm))
    return buf, tau, perm, nz


def householder_q(packed, tau):
    r, c = packed.shape
    Q = np.zeros((r, r), order="F")
    p = np.asfortranarray(packed)
    lib().orc_householder_q(_d(p), r, c, _d(_f64(tau)), _d(Q))
    return Q


def block_t_factor(V, tau):
    rows, n = V.shape
    T = np.zeros((n, n), order="F")
    Vf = np.asfortranarray(V)
    lib().orc_block_t_factor(_d(Vf), rows, n, _d(_f64(tau)), _d(T))
    return T


def synth_blocks(seed, nb, r, c, block0=0, lo=0.5, hi=5.0):
    """nb blocks of r x c, flat block-major / col-major in the block (the device block-COO layout)."""
    out = np.empty(nb * r * c)
    lib().orc_synth_fill_blocks(C.c_uint64(seed), C.c_int64(block0), C.c_int64(nb), r, c, C.c_double(lo),
                                C.c_double(hi), _d(out))
    return out


# ------------------------------------------------------------------------------ sparse helper
class SparseOut:
    def __init__(self, rows, cols, outer, inner, val, row_major):
        self.rows, self.cols, self.outer, self.inner, self.val, self.row_major = rows, cols, outer, inner, val, row_major

    def toarray(self):
        M = np.zeros((self.rows, self.cols))
        for o in range(len(self.outer) - 1):
            for p in range(self.outer[o], self.outer[o + 1]):
                if self.row_major:
                    M[o, self.inner[p]] += self.val[p]
                else:
                    M[self.inner[p], o] += self.val[p]
        return M

    def tocsc(self):
        import scipy.sparse as sp
        cls = sp.csr_matrix if self.row_major else sp.csc_matrix
        return cls((self.val, self.inner, self.outer), shape=(self.rows, self.cols))


# ------------------------------------------------------------------------------ block diagonal
class BlockDiagonalOracle:
    """Reference-faithful BlockDiagonalSparseQR (BlockDiagonalSparseQR.h:415-547, 258-280)."""

    def __init__(self, br, bc, values, n_rows=None, n_cols=None, colpiv=True, qformat=0, build_q=True):
        br, bc = _i32(br), _i32(bc)
        self.nb = len(br)
        self.rows = int(br.sum()) if n_rows is None else n_rows
        self.cols = int(bc.sum()) if n_cols is None else n_cols
        self.br, self.bc = br, bc
        values = _f64(values)
        self._h = C.c_void_p(lib().orc_bd_factorize(self.nb, _i(br), _i(bc), _d(values), self.rows, self.cols,
                                                     int(colpiv), int(qformat), int(build_q)))
        L = lib()
        self.info = L.orc_bd_info(self._h)
        self.rank = L.orc_bd_rank(self._h)
        self._nvals = len(values)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_bd_free(self._h)
            self._h = None

    def matrixQ(self):
        nnz = lib().orc_bd_q_nnz(self._h)
        outer = np.zeros(self.rows + 1, dtype=np.int32)
        inner = np.zeros(nnz, dtype=np.int32)
        val = np.zeros(nnz)
        lib().orc_bd_get_q(self._h, _i(outer), _i(inner), _d(val))
        return SparseOut(self.rows, self.rows, outer, inner, val, True)

    def matrixR(self):
        nnz = lib().orc_bd_r_nnz(self._h)
        outer = np.zeros(self.cols + 1, dtype=np.int32)
        inner = np.zeros(nnz, dtype=np.int32)
        val = np.zeros(nnz)
        lib().orc_bd_get_r(self._h, _i(outer), _i(inner), _d(val))
        return SparseOut(self.rows, self.cols, outer, inner, val, False)

    def colsPermutation(self):
        p = np.zeros(self.cols, dtype=np.int32)
        lib().orc_bd_get_perms(self._h, _i(p), None)
        return p

    def rowsPermutation(self):
        p = np.zeros(self.cols, dtype=np.int32)
        q = np.zeros(self.rows, dtype=np.int32)
        lib().orc_bd_get_perms(self._h, _i(p), _i(q))
        return q

    def packed(self):
        pk = np.zeros(self._nvals)
        tau = np.zeros(self.cols)
        lib().orc_bd_get_packed(self._h, _d(pk), _d(tau))
        return pk, tau

    def solve(self, b):
        b = _f64(b)
        x = np.zeros(self.cols)
        lib().orc_bd_solve(self._h, _d(b), _d(x))
        return x

    def apply_qt(self, b):
        b = _f64(b)
        y = np.zeros(self.rows)
        lib().orc_bd_apply_qt(self._h, _d(b), _d(y))
        return y

    def apply_q(self, b):
        b = _f64(b)
        y = np.zeros(self.rows)
        lib().orc_bd_apply_q(self._h, _d(b), _d(y))
        return y


def bd_compact_uniform(nb, r, c, values, b, colpiv=False, threads=1):
    """CPU-baseline variant B/C: packed reflectors + fused Q^T b + back substitution (OpenMP when threads>1).
    Returns dict(packed, tau, perm, x, seconds)."""
    vals = _f64(values).copy()
    b = None if b is None else _f64(b)
    x = np.zeros(nb * c)
    tau = np.zeros(nb * c)
    perm = np.zeros(nb * c, dtype=np.int32)
    sec = lib().orc_bd_compact_uniform(C.c_int64(nb), r, c, _d(vals), None if b is None else _d(b), _d(x), _d(tau),
                                       _i(perm), int(colpiv), int(threads))
    return dict(packed=vals, tau=tau, perm=perm, x=x, seconds=sec)


def bd_reference_uniform(nb, r, c, values, b, colpiv=False):
    """CPU-baseline variant A: the reference's explicit-Q / sparse-assembly algorithm, single thread."""
    values, b = _f64(values), _f64(b)
    x = np.zeros(nb * c)
    sec = lib().orc_bd_reference_uniform(C.c_int64(nb), r, c, _d(values), _d(b), _d(x), int(colpiv))
    return dict(x=x, seconds=sec)


def hardware_threads():
    return lib().orc_hardware_threads()


# ------------------------------------------------------------------------------ block structure
def block_banded_pattern(mat_rows, mat_cols, block_rows, block_cols, overlap, suggested_block_cols):
    cap = mat_cols + 8
    out = np.zeros(4 * cap, dtype=np.int32)
    n = lib().orc_block_banded_pattern(mat_rows, mat_cols, block_rows, block_cols, overlap, suggested_block_cols,
                                       _i(out), cap)
    return out[: 4 * n].reshape(n, 4).copy()  # idxRow, idxCol, numRows, numCols


# ------------------------------------------------------------------------------ banded
class BandedOracle:
    """Reference-faithful BandedBlockedSparseQR (BandedBlockedSparseQR.h:443-519), rows already permuted."""

    def __init__(self, A_csc, blocks):
        A = A_csc.tocsc()
        A.sort_indices()
        self.rows, self.cols = A.shape
        blocks = _i32(blocks)
        self._h = C.c_void_p(lib().orc_banded_factorize(self.rows, self.cols, _i(_i32(A.indptr)), _i(_i32(A.indices)),
                                                         _d(_f64(A.data)), len(blocks), _i(blocks)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_banded_free(self._h)
            self._h = None

    def matrixR(self):
        nnz = lib().orc_banded_r_nnz(self._h)
        outer = np.zeros(self.cols + 1, dtype=np.int32)
        inner = np.zeros(nnz, dtype=np.int32)
        val = np.zeros(nnz)
        lib().orc_banded_get_r(self._h, _i(outer), _i(inner), _d(val))
        return SparseOut(self.rows, self.cols, outer, inner, val, False)

    def blocks(self):
        out = []
        for k in range(lib().orc_banded_num_blocks(self._h)):
            d = np.zeros(5, dtype=np.int32)
            lib().orc_banded_block_dims(self._h, k, _i(d))
            Y = np.zeros((d[0], d[1]), order="F")
            T = np.zeros((d[1], d[1]), order="F")
            lib().orc_banded_get_block(self._h, k, _d(Y), _d(T))
            out.append(dict(Y=Y, T=T, row=int(d[2]), col=int(d[3]), numZeros=int(d[4])))
        return out

    def apply_q(self, v, transpose):
        v = _f64(v).copy()
        lib().orc_banded_apply_q(self._h, _d(v), int(transpose))
        return v

    def solve(self, b):
        b = _f64(b)
        x = np.zeros(self.cols)
        lib().orc_banded_solve(self._h, _d(b), _d(x))
        return x


# ------------------------------------------------------------------------------ block angular
class BlockAngularOracle:
    """Reference-faithful BlockAngularSparseQR (BlockAngularSparseQR.h:459-514) with a dense border.
    right_kind: 0 = ColPivHouseholderQR<MatrixXd>, 1 = BlockedThinDenseQR<MatrixXd, panel>,
    2 = BlockedThinSparseQR<SparseMatrix, panel> on a dense border (per-panel ColPiv, zero-pivot columns deferred)."""

    def __init__(self, J2, *, br=None, bc=None, values=None, left_colpiv=True, A_csc=None, blocks=None,
                 right_kind=0, panel=2):
        J2 = np.asfortranarray(J2, dtype=np.float64)
        self.rows, self.m2 = J2.shape
        if A_csc is None:
            br, bc = _i32(br), _i32(bc)
            self.m1 = int(bc.sum())
            values = _f64(values)
            self._h = C.c_void_p(lib().orc_angular_factorize_bd(len(br), _i(br), _i(bc), _d(values), self.rows,
                                                                 self.m1, _d(J2), self.m2, int(left_colpiv),
                                                                 right_kind, panel))
        else:
            A = A_csc.tocsc()
            A.sort_indices()
            self.m1 = A.shape[1]
            blocks = _i32(blocks)
            self._h = C.c_void_p(lib().orc_angular_factorize_banded(self.rows, self.m1, _i(_i32(A.indptr)),
                                                                     _i(_i32(A.indices)), _d(_f64(A.data)),
                                                                     len(blocks), _i(blocks), _d(J2), self.m2,
                                                                     right_kind, panel))
        self.cols = self.m1 + self.m2
        self.rank = lib().orc_angular_rank(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_angular_free(self._h)
            self._h = None

    def matrixR(self):
        nnz = lib().orc_angular_r_nnz(self._h)
        outer = np.zeros(self.cols + 1, dtype=np.int32)
        inner = np.zeros(nnz, dtype=np.int32)
        val = np.zeros(nnz)
        lib().orc_angular_get_r(self._h, _i(outer), _i(inner), _d(val))
        return SparseOut(self.rows, self.cols, outer, inner, val, False)

    def colsPermutation(self):
        p = np.zeros(self.cols, dtype=np.int32)
        lib().orc_angular_get_perms(self._h, _i(p), None)
        return p

    def rightPermutation(self):
        p = np.zeros(self.m2, dtype=np.int32)
        lib().orc_angular_get_right_perm(self._h, _i(p))
        return p

    def apply_qt(self, v):
        v = _f64(v)
        y = np.zeros(self.rows)
        lib().orc_angular_apply_qt(self._h, _d(v), _d(y))
        return y

    def solve(self, b):
        b = _f64(b)
        x = np.zeros(self.cols)
        lib().orc_angular_solve(self._h, _d(b), _d(x))
        return x


def angular_reference_uniform(nb, r, c, values, J2, b, left_colpiv=True):
    values, b = _f64(values), _f64(b)
    J2 = np.asfortranarray(J2, dtype=np.float64)
    x = np.zeros(nb * c + J2.shape[1])
    sec = lib().orc_angular_reference_uniform(C.c_int64(nb), r, c, _d(values), _d(J2), J2.shape[1], _d(b), _d(x),
                                              int(left_colpiv))
    return dict(x=x, seconds=sec)
