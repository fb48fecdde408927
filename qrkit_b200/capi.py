"""ctypes binding of the C ABI (include/qrkit_b200.h) — the same entry points a cgo/JNI/C++ host binds.

Used by the Python mirror of the reference's solver interface (qrkit_b200/solvers.py), tests/ and
bench.py.  Loading fails loudly when the CUDA library has not been built: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QRKIT_B200_LIB") or os.path.join(_PKG, "lib", "libqrkit_b200.so")   # the override is a development switch (tools/build_variant.py)
HEADER = os.path.join(os.path.dirname(_PKG), "include", "qrkit_b200.h")

QRK_STATUS_OK = 0
QRK_STATUS_INVALID_ARGUMENT = 1
QRK_STATUS_NOT_FACTORIZED = 2
QRK_STATUS_CUDA_ERROR = 3
QRK_STATUS_NO_DEVICE = 4
QRK_STATUS_ALLOC_FAILED = 5
QRK_STATUS_UNSUPPORTED = 6
QRK_STATUS_PEER_TIMEOUT = 7

QRK_INFO_SUCCESS, QRK_INFO_NUMERICAL_ISSUE, QRK_INFO_NO_CONVERGENCE, QRK_INFO_INVALID_INPUT = 0, 1, 2, 3
QRK_BLOCK_DIAGONAL, QRK_BLOCK_ANGULAR, QRK_BANDED_BLOCKED = 0, 1, 2
QRK_RIGHT_COLPIV, QRK_RIGHT_UNPIVOTED = 0, 1
QRK_PIVOT_NONE, QRK_PIVOT_COLPIV = 0, 1
QRK_FULL_Q, QRK_BLOCK_DIAGONAL_Q = 0, 1
QRK_HOST, QRK_DEVICE = 0, 1


class QrkDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("device", C.c_int32), ("num_blocks", C.c_int64),
        ("block_rows", C.c_int32), ("block_cols", C.c_int32),
        ("rows", C.POINTER(C.c_int32)), ("cols", C.POINTER(C.c_int32)),
        ("n_rows", C.c_int64), ("n_cols", C.c_int64),
        ("pivoting", C.c_int32), ("q_format", C.c_int32),
        ("border_cols", C.c_int32), ("block_overlap", C.c_int32),
        ("right_solver", C.c_int32), ("left_solver", C.c_int32), ("reserved", C.c_int32 * 2),
    ]


class QrkError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"qrkit_b200 status {status}: {msg}")
        self.status = status


_lib = None


def declared_symbols():
    """Every function the header declares (QRK_API <ret> name(...))."""
    with open(HEADER) as f:
        text = f.read()
    return re.findall(r"QRK_API\s+[\w\s\*]+?\b(qrk_\w+)\s*\(", text)


def lib():
    """Load libqrkit_b200.so.  Raises if it is missing — the product path never falls back to the CPU."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(qrkit_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.qrk_status_string.restype = C.c_char_p
        L.qrk_last_error.restype = C.c_char_p
        L.qrk_last_error.argtypes = [C.c_void_p]
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        sig = {
            "qrk_create": [C.POINTER(QrkDesc), C.POINTER(vp)],
            "qrk_destroy": [vp], "qrk_set_stream": [vp, vp], "qrk_synchronize": [vp],
            "qrk_set_blocks": [vp, vp, C.c_int], "qrk_adopt_blocks": [vp, vp],
            "qrk_total_values": [vp, C.POINTER(i64)],
            "qrk_analyze_pattern": [vp, vp], "qrk_factorize": [vp], "qrk_compute": [vp, vp, C.c_int],
            "qrk_compute_solve": [vp, vp, vp, vp, C.c_int], "qrk_factorize_solve": [vp, vp, vp, C.c_int],
            "qrk_rows": [vp, C.POINTER(i64)], "qrk_cols": [vp, C.POINTER(i64)], "qrk_rank": [vp, C.POINTER(i64)],
            "qrk_info": [vp, C.POINTER(i32)],
            "qrk_cols_permutation": [vp, vp, C.c_int], "qrk_rows_permutation": [vp, vp, C.c_int],
            "qrk_matrix_r_nnz": [vp, C.POINTER(i64)], "qrk_matrix_r": [vp, vp, vp, vp, C.c_int],
            "qrk_matrix_q_nnz": [vp, C.POINTER(i64)], "qrk_matrix_q": [vp, vp, vp, vp, C.c_int],
            "qrk_packed_factors": [vp, vp, vp, C.c_int],
            "qrk_apply_qt": [vp, vp, i64, vp, i64, i32, C.c_int],
            "qrk_apply_q": [vp, vp, i64, vp, i64, i32, C.c_int],
            "qrk_apply_qt_thin": [vp, vp, i64, vp, i64, i32, C.c_int],
            "qrk_apply_q_thin": [vp, vp, i64, vp, i64, i32, C.c_int],
            "qrk_solve": [vp, vp, i64, vp, i64, i32, C.c_int],
            "qrk_launch_count": [vp, C.POINTER(i64)],
            "qrk_set_border": [vp, vp, i64, C.c_int], "qrk_angular_set_world": [vp, i32],
            "qrk_angular_triangle_size": [vp, C.POINTER(i64)], "qrk_angular_local_triangle": [vp, vp, C.c_int],
            "qrk_angular_merge": [vp, vp, i32, C.c_int],
            "qrk_angular_xchg_buffer": [vp, C.POINTER(vp), C.POINTER(i64)],
            "qrk_angular_p2p_attach": [vp, C.POINTER(vp), i32, i32],
            "qrk_angular_p2p_status": [vp, C.POINTER(i32)], "qrk_angular_p2p_set_timeout": [vp, C.c_double],
            "qrk_enable_peer_access": [i32, i32],
            "qrk_ipc_export": [vp, vp], "qrk_ipc_import": [vp, C.POINTER(vp)], "qrk_ipc_close": [vp],
            "qrk_bind_host_thread_to_device": [i32, C.POINTER(i32), C.POINTER(i32)],
            "qrk_host_alloc": [C.POINTER(vp), i64, i32], "qrk_host_free": [vp],
            "qrk_synth_fill": [vp, C.c_uint64, i64, i64, i32, i32, C.c_double, C.c_double, vp],
            "qrk_device_count": [C.POINTER(C.c_int)],
            "qrk_ellipse_points": [vp, vp, i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, vp],
            "qrk_ellipse_assemble": [vp, vp, vp, i64, vp, vp, vp, vp, vp],
            "qrk_order_as_banded_as_possible": [i64, i64, vp, vp, vp, C.POINTER(i32)],
            "qrk_order_column_density": [i64, vp, vp],
            "qrk_detect_blocks": [i64, i64, vp, vp, i32, vp, i64, C.POINTER(i64), C.POINTER(i64)],
            "qrk_detect_band_starts": [i64, i64, vp, vp, vp, i64, C.POINTER(i64)],
            "qrk_create_banded_general": [vp, i64, i64, i64, i32, i32, C.POINTER(vp)],
            "qrk_block_diagonal_pattern": [i64, i64, i32, i32, vp, i64, C.POINTER(i64)],
            "qrk_block_banded_pattern": [i64, i64, i32, i32, i32, i32, vp, i64, C.POINTER(i64)],
            "qrk_extract_blocks": [i64, i64, vp, vp, vp, vp, vp, i64, vp],
        }
        for name, argtypes in sig.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        _lib = L
    return _lib


def check(status, handle=None):
    if status != QRK_STATUS_OK:
        L = lib()
        msg = L.qrk_status_string(status).decode()
        if handle:
            detail = L.qrk_last_error(handle).decode()
            if detail:
                msg += " — " + detail
        raise QrkError(status, msg)


def device_count() -> int:
    n = C.c_int(0)
    lib().qrk_device_count(C.byref(n))
    return n.value
