"""qrkit_b200 — B200-native (sm_100a) implementation of QRKit's structured-sparse QR hot path.

Product = qrkit_b200/csrc (CUDA kernels + C ABI, include/qrkit_b200.h) and the C++ façade under
include/qrkit_b200/.  The Python modules here only bind the C ABI for tests and benchmarks."""
from .capi import (QRK_BLOCK_DIAGONAL_Q, QRK_DEVICE, QRK_FULL_Q, QRK_HOST, QRK_INFO_INVALID_INPUT, QRK_INFO_SUCCESS,
                   QRK_PIVOT_COLPIV, QRK_PIVOT_NONE, QrkError, device_count)
from .solvers import BandedBlockedSparseQR, BandedSlabs, BlockAngularSparseQR, BlockDiagonalSparseQR, BlockMatrix1x2, SparseBlockDiagonal

__all__ = ["BandedBlockedSparseQR", "BandedSlabs", "BlockDiagonalSparseQR", "BlockAngularSparseQR", "BlockMatrix1x2", "SparseBlockDiagonal", "QrkError", "device_count", "QRK_PIVOT_COLPIV", "QRK_PIVOT_NONE",
           "QRK_FULL_Q", "QRK_BLOCK_DIAGONAL_Q", "QRK_HOST", "QRK_DEVICE", "QRK_INFO_SUCCESS", "QRK_INFO_INVALID_INPUT"]
