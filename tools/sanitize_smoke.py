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
# two-phase banded path: several groups (QRK_BANDED_GROUP), Q^T b and Q v on the stored factors
os.environ["QRK_BANDED_GROUP"] = "3"
for (br, bc, ov, nbk) in [(16, 24, 16, 11), (7, 4, 2, 9), (7, 2, 0, 5), (12, 8, 4, 10)]:
    slabs = uniform_blocks(nbk, br, bc)
    n_rows, n_cols = nbk * br, (nbk - 1) * (bc - ov) + bc
    s = qk.BandedBlockedSparseQR(slabs, num_blocks=nbk, block_rows=br, block_cols=bc, overlap=ov)
    bvec = vector(n_rows, seed=3)
    y = s.applyQtThin(bvec)                  # (the two-phase path keeps the thin Q; the exact n x n Q is the general chain's)
    print("banded2", br, bc, ov, float(np.abs(s.solve(bvec)).max()), float(np.abs(s.applyQThin(y)).max()), s.matrixR().toarray().shape)
del os.environ["QRK_BANDED_GROUP"]
# dense border: blocked compact-WY first stage (cluster panel + DMMA update) and the cluster-resident ColPiv second stage,
# ragged last panel (m2 = 21), unpivoted right solver, solve on the stored factors
for (nb, r, c, m2, right) in [(30, 7, 2, 21, 0), (30, 7, 2, 24, 1), (120, 7, 2, 40, 0)]:
    vals = uniform_blocks(nb, r, c); J2 = dense_border(nb * r, m2); b = vector(nb * r, seed=5)
    mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), J2)
    s = qk.BlockAngularSparseQR(mat, pivoting=1, right_solver=right)
    print("wide blocked", m2, right, float(np.abs(s.solve(b)).max()), s.rank(),
          float(np.abs(qk.BlockAngularSparseQR(pivoting=1, right_solver=right).compute_solve(mat, b)).max()))
# device-side ellipse assembly
import ctypes as C
import torch
from qrkit_b200 import capi
L = capi.lib()
n = 1000
px = torch.empty(n, dtype=torch.float64, device="cuda"); py = torch.empty_like(px)
capi.check(L.qrk_ellipse_points(px.data_ptr(), py.data_ptr(), n, 7.5, 2.0, 17.0, 23.0, 0.23, None))
params = torch.cat([torch.arange(n, dtype=torch.float64, device="cuda") * (1.3 * np.pi / n), torch.tensor([7.0, 2.5, 17.0, 23.0, 0.0], dtype=torch.float64, device="cuda")])
J1 = torch.empty(2 * n, dtype=torch.float64, device="cuda"); J2d = torch.empty(10 * n, dtype=torch.float64, device="cuda")
rhs = torch.empty(2 * n, dtype=torch.float64, device="cuda"); cost = torch.zeros(1, dtype=torch.float64, device="cuda")
capi.check(L.qrk_ellipse_assemble(px.data_ptr(), py.data_ptr(), params.data_ptr(), n, J1.data_ptr(), J2d.data_ptr(), rhs.data_ptr(), cost.data_ptr(), None))
torch.cuda.synchronize()
print("ellipse", float(cost.item()))
# round 2: the 2 x 1 block-angular kernels (unstaged K1, one-wave K3 with private cp.async slots), ragged point counts that
# leave partial iterations / batches, ColPiv and unpivoted left blocks, compute() + solve() (stored Abot panel), full Q products
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import ellipse_problem
for npts in (5, 129, 1000, 4099):
    J1v, J2e, rhse = ellipse_problem(npts)
    mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(J1v, block_rows=2, block_cols=1), J2e)
    for piv in (0, 1):
        s = qk.BlockAngularSparseQR(pivoting=piv)
        xf = s.compute_solve(mat, rhse)
        s2 = qk.BlockAngularSparseQR(mat, pivoting=piv)
        xs = s2.solve(rhse)
        print("angular 2x1", npts, piv, float(np.abs(xf - xs).max()), s2.rank())
    y = s2.applyQt(rhse)
    print("angular full Q", npts, float(np.abs(s2.applyQ(y) - rhse).max()))
# general banded chain (run-time window table) and the thin-sparse right solver
os.environ["QRK_BANDED_GENERIC"] = "1"
slabs = uniform_blocks(9, 9, 5)
s = qk.BandedBlockedSparseQR(slabs, num_blocks=9, block_rows=9, block_cols=5, overlap=2)
print("banded generic", float(np.abs(s.solve(vector(81, seed=3))).max()))
del os.environ["QRK_BANDED_GENERIC"]
nb, r, c, m2 = 40, 7, 2, 20
vals = uniform_blocks(nb, r, c); J2 = dense_border(nb * r, m2); J2[:, 3] = 0.0; b = vector(nb * r, seed=5)
mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), J2)
s = qk.BlockAngularSparseQR(pivoting=1, right_solver=2)
print("thin sparse right solver", float(np.abs(s.compute_solve(mat, b)).max()), s.rank())     # (a deferred zero column: fused path)
# register-resident ColPiv triangle kernel (dense_tri_reg.cuh: borders of 65..384 columns), ragged width, two zero columns
nb, r, c, m2 = 30, 7, 2, 70
vals = uniform_blocks(nb, r, c); J2 = dense_border(nb * r, m2); J2[:, 11] = 0.0; J2[:, 40] = 0.0; b = vector(nb * r, seed=5)
mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), J2)
s = qk.BlockAngularSparseQR(mat, pivoting=1)
print("register-resident triangle", s.rank(), float(np.abs(s.solve(b)).max()))
# steps replayed from library-captured CUDA graphs: three calls on one handle (eager, capture + launch, replay)
nb, r, c, m2 = 30, 7, 2, 24
vals = uniform_blocks(nb, r, c); J2 = dense_border(nb * r, m2); b = vector(nb * r, seed=5)
mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), J2)
s = qk.BlockAngularSparseQR(pivoting=1)
print("wide step graph", [float(np.abs(s.compute_solve(mat, b)).max()) for _ in range(3)])
J1v, J2e, rhse = ellipse_problem(777)
mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(J1v, block_rows=2, block_cols=1), J2e)
s = qk.BlockAngularSparseQR(pivoting=0)
print("tsqr step graph", [float(np.abs(s.compute_solve(mat, rhse)).max()) for _ in range(3)])
