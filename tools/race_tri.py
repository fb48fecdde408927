"""racecheck target: the register-resident ColPiv triangle kernel alone (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import uniform_blocks, vector, dense_border
import qrkit_b200 as qk
nb, r, c, m2 = 30, 7, 2, 70
vals = uniform_blocks(nb, r, c); J2 = dense_border(nb * r, m2); b = vector(nb * r, seed=5)
mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), J2)
s = qk.BlockAngularSparseQR(mat, pivoting=1)
print(s.rank())
