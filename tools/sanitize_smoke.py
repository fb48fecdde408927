"""Small invocations of the newer kernels for compute-sanitizer (memcheck / racecheck / synccheck): development tool."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import uniform_blocks, vector, dense_border
import qrkit_b200 as qk

for (r, c) in [(32, 16), (48, 24), (64, 32), (100, 50), (128, 64), (40, 9)]:
    nb = 3
    vals = uniform_blocks(nb, r, c); b = vector(nb * r, seed=7)
    x = qk.BlockDiagonalSparseQR(pivoting=0).compute_solve(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), b)
    print("wy", r, c, float(np.abs(x).max()))
nb, r, c, m2 = 24, 7, 2, 12
vals = uniform_blocks(nb, r, c); J2 = dense_border(nb * r, m2); b = vector(nb * r, seed=5)
mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), J2)
s = qk.BlockAngularSparseQR(mat, pivoting=1)
print("wide", float(np.abs(s.solve(b)).max()), s.rank())
slabs = uniform_blocks(6, 16, 24)
print("banded", float(np.abs(qk.BandedBlockedSparseQR(block_rows=16, block_cols=24, overlap=16).compute_solve(slabs, vector(96, seed=3), 6)).max()))
