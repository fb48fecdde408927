"""Debug driver for the blocked-WY kernel: per-shape parity vs the oracle (test infrastructure, not product)."""
import sys, os, faulthandler
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import uniform_blocks, vector, rel
from oracle import oracle as orc
import qrkit_b200 as qk

shapes = [(32, 16), (48, 24), (64, 32), (128, 64), (20, 12), (40, 9), (127, 63), (100, 100), (17, 16), (128, 128)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (r, c) in shapes:
    nb = 5
    vals = uniform_blocks(nb, r, c)
    b = vector(nb * r, seed=7)
    mat = qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c)
    solver = qk.BlockDiagonalSparseQR(pivoting=0)
    print("shape", r, c, flush=True)
    x = solver.compute_solve(mat, b)
    ref = orc.bd_compact_uniform(nb, r, c, vals, b, colpiv=False)
    pk, tau = solver.packed()
    P = pk.reshape(nb, c, r); Pr = ref["packed"].reshape(nb, c, r)
    colerr = [rel(P[0, j], Pr[0, j]) for j in range(c)]
    print("  packed", rel(pk, ref["packed"]), "tau", rel(tau, ref["tau"]), "x", rel(x, ref["x"]), flush=True)
    print("  first bad col (block 0):", next((j for j, e in enumerate(colerr) if e > 1e-12), None), flush=True)
