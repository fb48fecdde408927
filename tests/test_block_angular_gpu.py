"""GPU parity tests of the block-angular path (left block diagonal + dense border, TSQR merge) against the CPU oracle's
restatement of BlockAngularSparseQR<BlockDiagonalSparseQR, ColPivHouseholderQR<MatrixXd>> (BlockAngularSparseQR.h:459-514).
Inputs: the ellipse-fit Jacobian of bench/bench_sparse_qr_extra.cpp:79-114 (BASELINE configs 1 and 3) and random blocks
with a dense random border (the shape of test/test-qrkit.cpp:135-165)."""
import time

import numpy as np
import pytest

from helpers import SEED_A, blocks_to_dense, dense_border, ellipse_problem, rel, uniform_blocks, vector

pytestmark = pytest.mark.gpu
TOL_R, TOL_X = 1e-12, 1e-10


@pytest.fixture(scope="module")
def qk():
    import qrkit_b200 as q
    if q.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests must run on the B200 box (there is no CPU fallback)")
    return q


def _sign_fix(R2, R2ref):
    s = np.sign(np.diag(R2)) * np.sign(np.diag(R2ref))
    s[s == 0] = 1.0
    return R2 * s[:, None]


def _check(qk, oracle, vals, r, c, J2, b, piv, tol_x=TOL_X):
    nb = len(vals) // (r * c)
    m1, m2 = nb * c, J2.shape[1]
    left = qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c)
    mat = qk.BlockMatrix1x2(left, J2)
    ref = oracle.BlockAngularOracle(J2, br=np.full(nb, r), bc=np.full(nb, c), values=vals, left_colpiv=bool(piv), right_kind=0)
    x_ref = ref.solve(b)
    # fused one-pass compute + solve
    s1 = qk.BlockAngularSparseQR(pivoting=piv)
    x1 = s1.compute_solve(mat, b)
    assert rel(x1, x_ref) <= tol_x
    # compute(), then accessors and solve()
    s2 = qk.BlockAngularSparseQR(mat, pivoting=piv)
    assert s2.rows() == nb * r and s2.cols() == m1 + m2
    assert s2.rank() == ref.rank and s2.info() == qk.QRK_INFO_SUCCESS
    P = s2.colsPermutation()
    assert np.array_equal(P, ref.colsPermutation()), "P_c = [P1; m1 + P2] must be bit-exact"
    R, Rr = s2.matrixR(), ref.matrixR()
    assert np.array_equal(R.outer, Rr.outer) and np.array_equal(R.inner, Rr.inner)
    Rd, Rrd = R.toarray(), Rr.toarray()
    assert rel(Rd[:m1, :], Rrd[:m1, :]) <= TOL_R                     # [R1, Atop*P2]: determined by Q1, sign included
    R2 = _sign_fix(Rd[m1:m1 + m2, m1:], Rrd[m1:m1 + m2, m1:])         # R2 equal up to per-row sign
    assert rel(R2, Rrd[m1:m1 + m2, m1:]) <= TOL_R * 10
    A = np.hstack([blocks_to_dense(vals, np.full(nb, r), np.full(nb, c)), J2])
    AP = A[:, P]
    Rt = Rd[:m1 + m2, :]
    assert rel(Rt.T @ Rt, AP.T @ AP) <= 1e-12                         # (AP)^T AP = R^T R  <=>  AP = QR with Q orthonormal
    assert rel(s2.solve(b), x_ref) <= tol_x
    x_true = vector(m1 + m2, seed=77)
    assert rel(s2.solve(A @ x_true), x_true) <= tol_x
    # matrixQ().transpose() * v = [I 0; 0 Q2^T] Q1^T v (BlockAngularSparseQR.h:607-625): thin part of the left factor, then the
    # right solver's Q2^T on the complement rows; its first m2 entries equal the reference's up to the row signs of R2
    y = s2.applyQt(b)
    yref = ref.apply_qt(b)
    assert rel(y[:m1], yref[:m1]) <= TOL_R
    assert rel(np.abs(y[m1:m1 + m2]), np.abs(yref[m1:m1 + m2])) <= 1e-10
    assert abs(np.linalg.norm(y) - np.linalg.norm(b)) <= 1e-13 * np.linalg.norm(b)      # Q is orthogonal on all n rows
    assert rel(s2.applyQ(y), b) <= 1e-13                                                  # Q Q^T = I
    if nb * r <= 4000:
        # the reference's own identities (test/test-qrkit.cpp:251-254) on the full n x n operator: Q^T (A P) = R, Q R = A P
        n = nb * r
        Rfull = np.zeros((n, m1 + m2)); Rfull[:m1 + m2, :] = Rt
        assert rel(s2.applyQt(np.asfortranarray(AP)), Rfull) <= 1e-13
        assert rel(s2.applyQ(np.asfortranarray(Rfull)), AP) <= 1e-13
    # the fused compute_solve() of the TSQR path keeps no residual panel: the Q2 stage cannot be formed, and the call says so
    # (the dense right-block path stores its factor either way)
    try:
        y1 = s1.applyQt(b)
    except qk.QrkError:
        assert m2 <= 8
    else:
        assert rel(y1, y) <= 1e-13


@pytest.mark.parametrize("n", [500, 2000, 10000])
@pytest.mark.parametrize("piv", [0, 1])
def test_ellipse_jacobian_vs_oracle(qk, oracle, n, piv):
    """BASELINE config 1: ellipse Jacobian, N blocks of 2x1 + 2N x 5 border, rhs = residual at the initial iterate."""
    J1, J2, rhs = ellipse_problem(n)
    _check(qk, oracle, J1, 2, 1, J2, rhs, piv, tol_x=TOL_X)       # kappa of the Schur-reduced border is 16: no slack needed


@pytest.mark.parametrize("r,c", [(2, 1), (3, 1), (4, 2), (7, 2)])
@pytest.mark.parametrize("m2", [1, 2, 3, 4, 5, 6, 7, 8])
def test_random_border_vs_oracle(qk, oracle, r, c, m2):
    nb = 333
    vals = uniform_blocks(nb, r, c)
    J2 = dense_border(nb * r, m2)
    b = vector(nb * r, seed=5)
    _check(qk, oracle, vals, r, c, J2, b, piv=1)


def test_multi_gpu_exchange_on_one_device(qk, oracle):
    """The N-GPU algorithm with the ranks emulated as handles on one device: contiguous block ranges per rank,
    per-rank TSQR triangles 'all-gathered' on the host, root merged redundantly by every rank."""
    n, world = 4000, 4
    J1, J2, rhs = ellipse_problem(n)
    ref = oracle.BlockAngularOracle(J2, br=np.full(n, 2), bc=np.full(n, 1), values=J1, left_colpiv=True, right_kind=0)
    x_ref = ref.solve(rhs)
    per = n // world
    solvers, tris = [], []
    for g in range(world):
        sl = slice(g * per, (g + 1) * per)
        rows = slice(2 * g * per, 2 * (g + 1) * per)
        mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(J1[2 * g * per:2 * (g + 1) * per].copy(), block_rows=2, block_cols=1),
                                J2[rows, :].copy())
        s = qk.BlockAngularSparseQR(world=world)
        s.compute_solve(mat, rhs[rows].copy())
        tris.append(s.local_triangle())
        solvers.append(s)
    gathered = np.concatenate(tris)
    x2s = []
    for g, s in enumerate(solvers):
        xg = s.merge(gathered)
        assert rel(xg[:per], x_ref[g * per:(g + 1) * per]) <= 1e-9
        assert rel(xg[per:], x_ref[n:]) <= 1e-9
        x2s.append(xg[per:].copy())
        assert np.array_equal(s.colsPermutation()[per:] - per, ref.rightPermutation())
    for g in range(1, world):
        assert np.array_equal(x2s[g], x2s[0]), "the redundantly computed shared parameters must be bit-identical on every rank"


@pytest.mark.parametrize("br,bc,ov,nb,m2", [(7, 4, 2, 60, 5), (7, 2, 0, 40, 9), (16, 24, 16, 30, 24), (12, 8, 4, 50, 3), (16, 24, 16, 70, 40)])
@pytest.mark.parametrize("right", [0, 1])
def test_banded_left_block_vs_lapack(qk, br, bc, ov, nb, m2, right):
    """BlockAngularSparseQR<BandedBlockedSparseQR, RightSolver> — the solver combination of the reference's own block-angular
    tests (test/test-qrkit.cpp:44-57) and of the QRkitBB benchmark column: J1 block banded, dense border.  Checked against
    LAPACK on the assembled matrix: x (lstsq), (AP)^T AP = R^T R, P2 = dgeqp3's permutation of the residual rows
    (Q1 complete from numpy), rank, and the left factor's Q^T through R1^T (Q1^T b) = J1^T b."""
    import scipy.linalg as sla
    from helpers import slabs_to_sparse
    slabs = uniform_blocks(nb, br, bc)
    A1 = slabs_to_sparse(slabs, nb, br, bc, ov).toarray()
    n, m1 = A1.shape
    J2 = dense_border(n, m2)
    b = vector(n, seed=5)
    A = np.hstack([A1, J2])
    x_ls = np.linalg.lstsq(A, b, rcond=None)[0]
    mat = qk.BlockMatrix1x2(qk.BandedSlabs(slabs, num_blocks=nb, block_rows=br, block_cols=bc, overlap=ov), J2)
    s = qk.BlockAngularSparseQR(mat, right_solver=right)
    assert s.rows() == n and s.cols() == m1 + m2 and s.rank() == m1 + m2 and s.info() == qk.QRK_INFO_SUCCESS
    P = s.colsPermutation()
    assert np.array_equal(P[:m1], np.arange(m1))
    if right == 0:
        Q1 = np.linalg.qr(A1, mode="complete")[0]
        Abot = (Q1.T @ J2)[m1:, :]
        assert np.array_equal(P[m1:] - m1, sla.qr(Abot, mode="r", pivoting=True)[1])
    else:
        assert np.array_equal(P[m1:], m1 + np.arange(m2))
    R = s.matrixR().toarray()[:m1 + m2, :]
    assert np.allclose(np.tril(R, -1), 0.0)
    AP = A[:, P]
    assert rel(R.T @ R, AP.T @ AP) <= 1e-12
    assert rel(s.solve(b), x_ls) <= 1e-10
    x_true = vector(m1 + m2, seed=77)
    assert rel(s.solve(A @ x_true), x_true) <= 1e-10
    s2 = qk.BlockAngularSparseQR(right_solver=right)
    assert rel(s2.compute_solve(mat, b), x_ls) <= 1e-10
    y = s.applyQtThin(b)                                        # [Q1thin^T b; z2]: the vector _solve_impl back-substitutes
    assert y.shape == (m1 + m2,)
    assert rel(R[:m1, :m1].T @ y[:m1], A1.T @ b) <= 1e-11
    assert rel(R.T @ y, AP.T @ b) <= 1e-10                      # R^T (Q_thin^T b) = (A P)^T b
    with pytest.raises(qk.QrkError):                            # no n x n Q for a banded left factor: refused, not approximated
        s.applyQt(b)


def test_reference_test4_with_its_banded_left_solver(qk):
    """The reference's block-angular test at its own sizes AND with its own solver combination (test/test-qrkit.cpp:44-48,
    388-391): 7168 rows, left block 1024 slabs of 7x2 (2048 columns) factored by BandedBlockedSparseQR, a fully dense border of
    384 columns factored by ColPivHouseholderQR.  Asserts what the reference asserts (x recovered from a consistent system,
    :289) at 1e-10, plus x of a least-squares right-hand side, rank and (AP)^T AP = R^T R."""
    from helpers import slabs_to_sparse
    nb, br, bc, ov, m2 = 1024, 7, 2, 0, 384
    slabs = uniform_blocks(nb, br, bc)
    A1 = slabs_to_sparse(slabs, nb, br, bc, ov).toarray()
    n, m1 = A1.shape
    J2 = dense_border(n, m2)
    A = np.hstack([A1, J2])
    mat = qk.BlockMatrix1x2(qk.BandedSlabs(slabs, num_blocks=nb, block_rows=br, block_cols=bc, overlap=ov), J2)
    s = qk.BlockAngularSparseQR(mat)
    assert s.rank() == m1 + m2 and s.info() == qk.QRK_INFO_SUCCESS
    x_true = vector(m1 + m2, seed=21)
    assert rel(s.solve(A @ x_true), x_true) <= 1e-10
    b = vector(n, seed=22)
    assert rel(s.solve(b), np.linalg.lstsq(A, b, rcond=None)[0]) <= 1e-10
    P = s.colsPermutation()
    R = s.matrixR().toarray()[:m1 + m2, :]
    AP = A[:, P]
    assert rel(R.T @ R, AP.T @ AP) <= 1e-12


def test_device_side_ellipse_assembly_and_gauss_newton(qk):
    """qrk_ellipse_assemble writes the functor's Jacobian (bench/bench_sparse_qr_extra.cpp:79-114) straight into
    the device buffers of the block-angular solver; a Gauss-Newton loop that never leaves the device recovers the ellipse."""
    import ctypes as C
    import torch
    from qrkit_b200 import capi
    from qrkit_b200.capi import QRK_DEVICE, QrkDesc, check
    L = capi.lib()
    n = 20000
    vp = lambda t: C.c_void_p(t.data_ptr())
    truth = np.array([7.5, 2.0, 17.0, 23.0, 0.23])
    px = torch.empty(n, dtype=torch.float64, device="cuda"); py = torch.empty_like(px)
    check(L.qrk_ellipse_points(vp(px), vp(py), n, *truth, None))
    J1r, J2r, rhsr = ellipse_problem(n)
    t0 = np.arange(n) * (1.3 * np.pi / n)
    pxh, pyh = px.cpu().numpy(), py.cpu().numpy()
    p0 = np.array([0.5 * (pxh.max() - pxh.min()), 0.5 * (pyh.max() - pyh.min()), 0.5 * (pxh.max() + pxh.min()), 0.5 * (pyh.max() + pyh.min()), 0.0])
    params = torch.from_numpy(np.concatenate([t0, p0])).cuda()
    J1 = torch.empty(2 * n, dtype=torch.float64, device="cuda")
    J2 = torch.empty(10 * n, dtype=torch.float64, device="cuda")
    rhs = torch.empty(2 * n, dtype=torch.float64, device="cuda")
    cost = torch.zeros(8, dtype=torch.float64, device="cuda")
    check(L.qrk_ellipse_assemble(vp(px), vp(py), vp(params), n, vp(J1), vp(J2), vp(rhs), vp(cost), None))
    torch.cuda.synchronize()
    assert rel(J1.cpu().numpy(), J1r) <= 1e-13                                   # same functor as the host generator
    assert rel(J2.cpu().numpy().reshape(5, 2 * n).T, J2r) <= 1e-13
    assert rel(-rhs.cpu().numpy(), rhsr) <= 1e-11
    assert abs(cost[0].item() - float(rhsr @ rhsr)) <= 1e-10 * float(rhsr @ rhsr)
    d = QrkDesc()
    d.kind, d.num_blocks, d.block_rows, d.block_cols, d.pivoting, d.border_cols = capi.QRK_BLOCK_ANGULAR, n, 2, 1, 1, 5
    h = C.c_void_p()
    check(L.qrk_create(C.byref(d), C.byref(h)))
    check(L.qrk_set_border(h, vp(J2), 2 * n, QRK_DEVICE), h)
    step = torch.empty(n + 5, dtype=torch.float64, device="cuda")
    for it in range(6):
        check(L.qrk_ellipse_assemble(vp(px), vp(py), vp(params), n, vp(J1), vp(J2), vp(rhs), C.c_void_p(cost.data_ptr() + 8 * (it + 1)), None))
        check(L.qrk_compute_solve(h, vp(J1), vp(rhs), vp(step), QRK_DEVICE), h)
        check(L.qrk_synchronize(h), h)
        params.add_(step)
        torch.cuda.synchronize()
    L.qrk_destroy(h)
    c = cost.cpu().numpy()
    assert c[6] < 1e-18 * c[1], c                                                # quadratic convergence to the exact fit
    assert np.abs(params[n:].cpu().numpy() - truth).max() <= 1e-9
    assert np.abs(params[:n].cpu().numpy() - t0).max() <= 1e-9


@pytest.mark.parametrize("piv", [0, 1])
def test_fused_peer_exchange_two_ranks_on_one_device(qk, oracle, piv):
    """qrk_angular_p2p_attach: the per-GPU triangles are exchanged inside the TSQR root kernel (stores into the peers'
    exchange buffers + flags) instead of all-gather + qrk_angular_merge.  Two emulated ranks = two handles on two streams
    of one device, their exchange buffers attached by plain pointer; three steps (both buffer parities) with different
    right-hand sides; x2 bit-identical on both ranks."""
    import ctypes as C
    import torch
    from qrkit_b200 import capi
    from qrkit_b200.capi import QRK_DEVICE, QrkDesc, check
    L = capi.lib()
    n, world = 3000, 2
    per = n // world
    J1, J2, rhs0 = ellipse_problem(n)
    ref = oracle.BlockAngularOracle(J2, br=np.full(n, 2), bc=np.full(n, 1), values=J1, left_colpiv=bool(piv), right_kind=0)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    vp = lambda t: C.c_void_p(t.data_ptr())
    hs, bufs, xbufs = [], [], []
    for g in range(world):
        rows = slice(2 * g * per, 2 * (g + 1) * per)
        dJ1, dJ2 = dev(J1[rows]), dev(J2[rows, :].T)
        d = QrkDesc()
        d.kind, d.num_blocks, d.block_rows, d.block_cols, d.pivoting, d.border_cols = capi.QRK_BLOCK_ANGULAR, per, 2, 1, piv, 5
        h = C.c_void_p()
        check(L.qrk_create(C.byref(d), C.byref(h)))
        check(L.qrk_angular_set_world(h, world), h)
        check(L.qrk_set_border(h, vp(dJ2), 2 * per, QRK_DEVICE), h)
        check(L.qrk_set_blocks(h, vp(dJ1), QRK_DEVICE), h)          # allocations happen before any kernel spins
        xb, nb_ = C.c_void_p(), C.c_int64()
        check(L.qrk_angular_xchg_buffer(h, C.byref(xb), C.byref(nb_)), h)
        hs.append(h); bufs.append((dJ1, dJ2)); xbufs.append(xb.value)
    peers = (C.c_void_p * world)(*xbufs)
    for g, h in enumerate(hs):
        check(L.qrk_angular_p2p_attach(h, peers, world, g), h)
    for step in range(3):
        rhs = rhs0 if step == 0 else vector(2 * n, seed=100 + step)
        x_ref = ref.solve(rhs)
        dbs = [dev(rhs[2 * g * per:2 * (g + 1) * per]) for g in range(world)]
        dxs = [torch.zeros(per + 5, dtype=torch.float64, device="cuda") for _ in range(world)]
        for g, h in enumerate(hs):                                 # asynchronous: rank 0's root kernel waits for rank 1's
            check(L.qrk_compute_solve(h, vp(bufs[g][0]), vp(dbs[g]), vp(dxs[g]), QRK_DEVICE), h)
        x2s = []
        for g, h in enumerate(hs):
            to = C.c_int32(-1)
            check(L.qrk_angular_p2p_status(h, C.byref(to)), h)
            assert to.value == 0, "peer exchange timed out"
            xg = dxs[g].cpu().numpy()
            assert rel(xg[:per], x_ref[g * per:(g + 1) * per]) <= 1e-9
            assert rel(xg[per:], x_ref[n:]) <= 1e-9
            x2s.append(xg[per:].copy())
        assert np.array_equal(x2s[0], x2s[1])
    for h in hs:
        L.qrk_destroy(h)


@pytest.mark.parametrize("piv", [0, 1])
def test_multi_gpu_exchange_with_device_pointers(qk, oracle, piv):
    """The exchange as bench_extra.py / an NCCL caller runs it: every buffer (blocks, border, rhs, x, the local triangle
    and the gathered triangles) is a DEVICE pointer; two emulated ranks on one device."""
    import ctypes as C
    import torch
    from qrkit_b200 import capi
    from qrkit_b200.capi import QRK_DEVICE, QrkDesc, check
    L = capi.lib()
    n, world = 3000, 2
    J1, J2, rhs = ellipse_problem(n)
    ref = oracle.BlockAngularOracle(J2, br=np.full(n, 2), bc=np.full(n, 1), values=J1, left_colpiv=bool(piv), right_kind=0)
    x_ref = ref.solve(rhs)
    per = n // world
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    vp = lambda t: C.c_void_p(t.data_ptr())
    hs, bufs, tris = [], [], []
    for g in range(world):
        rows = slice(2 * g * per, 2 * (g + 1) * per)
        dJ1, dJ2, db = dev(J1[rows]), dev(J2[rows, :].T), dev(rhs[rows])     # border column-major 2*per x 5
        dx = torch.zeros(per + 5, dtype=torch.float64, device="cuda")
        d = QrkDesc()
        d.kind, d.num_blocks, d.block_rows, d.block_cols, d.pivoting, d.border_cols = capi.QRK_BLOCK_ANGULAR, per, 2, 1, piv, 5
        h = C.c_void_p()
        check(L.qrk_create(C.byref(d), C.byref(h)))
        check(L.qrk_angular_set_world(h, world), h)
        check(L.qrk_set_border(h, vp(dJ2), 2 * per, QRK_DEVICE), h)
        check(L.qrk_compute_solve(h, vp(dJ1), vp(db), vp(dx), QRK_DEVICE), h)
        tsz = C.c_int64()
        check(L.qrk_angular_triangle_size(h, C.byref(tsz)), h)
        tri = torch.empty(tsz.value, dtype=torch.float64, device="cuda")
        check(L.qrk_angular_local_triangle(h, vp(tri), QRK_DEVICE), h)
        check(L.qrk_synchronize(h), h)
        hs.append(h); bufs.append((dJ1, dJ2, db, dx)); tris.append(tri)
    gathered = torch.cat(tris)
    x2s = []
    for g, h in enumerate(hs):
        check(L.qrk_angular_merge(h, vp(gathered), world, QRK_DEVICE), h)
        check(L.qrk_synchronize(h), h)
        xg = bufs[g][3].cpu().numpy()
        assert rel(xg[:per], x_ref[g * per:(g + 1) * per]) <= 1e-9
        assert rel(xg[per:], x_ref[n:]) <= 1e-9
        x2s.append(xg[per:].copy())
        L.qrk_destroy(h)
    assert np.array_equal(x2s[0], x2s[1])


@pytest.mark.parametrize("r,c,m2,nb", [(2, 1, 9, 200), (7, 2, 24, 120), (8, 4, 5, 150), (7, 2, 96, 96), (16, 8, 40, 40), (7, 2, 21, 30), (7, 2, 19, 900)])
@pytest.mark.parametrize("piv", [0, 1])
def test_wide_border_vs_oracle(qk, oracle, r, c, m2, nb, piv):
    """Borders wider than the in-SM TSQR path (m2 > 8) and left blocks outside its shape list take the dense right-block
    solver (dense_border.cuh: Eigen's ColPivHouseholderQR on the residual rows): same checks, P2 and R2 included."""
    vals = uniform_blocks(nb, r, c)
    J2 = dense_border(nb * r, m2)
    b = vector(nb * r, seed=5)
    _check(qk, oracle, vals, r, c, J2, b, piv=piv)


def test_reference_test4_border_384(qk, oracle):
    """The reference's block-angular test sizes (test/test-qrkit.cpp:388-391): 7168 rows, 2048 left columns as 1024 blocks
    of 7x2, a fully dense border of 384 columns, right solver ColPivHouseholderQR<MatrixXd> (:46-48).  The test asserts what
    the reference asserts (x recovered from a consistent system, :289) at 1e-10 instead of 1e-6, plus P2 and rank."""
    nb, r, c, m2 = 1024, 7, 2, 384
    vals = uniform_blocks(nb, r, c)
    J2 = dense_border(nb * r, m2)
    A = np.hstack([blocks_to_dense(vals, np.full(nb, r), np.full(nb, c)), J2])
    x_true = vector(nb * c + m2, seed=21)
    mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), J2)
    s = qk.BlockAngularSparseQR(mat, pivoting=1)
    assert s.rank() == nb * c + m2 and s.info() == qk.QRK_INFO_SUCCESS
    assert rel(s.solve(A @ x_true), x_true) <= 1e-10
    ref = oracle.BlockAngularOracle(J2, br=np.full(nb, r), bc=np.full(nb, c), values=vals, left_colpiv=True, right_kind=0)
    assert np.array_equal(s.colsPermutation(), ref.colsPermutation())
    b = vector(nb * r, seed=22)
    assert rel(s.solve(b), ref.solve(b)) <= 1e-10
    s2 = qk.BlockAngularSparseQR(pivoting=1)
    assert rel(s2.compute_solve(mat, b), ref.solve(b)) <= 1e-10


@pytest.mark.parametrize("r,c,m2,nb", [(7, 2, 5, 100), (7, 2, 48, 64), (4, 2, 9, 128), (7, 2, 20, 40), (7, 2, 35, 900)])
def test_unpivoted_right_solver_vs_oracle(qk, oracle, r, c, m2, nb):
    """RightSolver = BlockedThinDenseQR<MatrixXd, 2> (reference test 5, test/test-qrkit.cpp:53-56, 294-327): no column
    pivoting in the right block, P2 = identity, rank = cols; R2 is unique up to row signs whatever the panel width."""
    vals = uniform_blocks(nb, r, c)
    J2 = dense_border(nb * r, m2)
    b = vector(nb * r, seed=6)
    m1 = nb * c
    ref = oracle.BlockAngularOracle(J2, br=np.full(nb, r), bc=np.full(nb, c), values=vals, left_colpiv=True, right_kind=1, panel=2)
    mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), J2)
    s = qk.BlockAngularSparseQR(mat, pivoting=1, right_solver=1)
    assert s.rank() == ref.rank == m1 + m2
    assert np.array_equal(s.colsPermutation(), ref.colsPermutation())
    assert np.array_equal(s.colsPermutation()[m1:], m1 + np.arange(m2))
    Rd, Rrd = s.matrixR().toarray(), ref.matrixR().toarray()
    assert rel(Rd[:m1, :], Rrd[:m1, :]) <= TOL_R
    assert rel(_sign_fix(Rd[m1:m1 + m2, m1:], Rrd[m1:m1 + m2, m1:]), Rrd[m1:m1 + m2, m1:]) <= TOL_R * 10
    assert rel(s.solve(b), ref.solve(b)) <= TOL_X
    assert rel(qk.BlockAngularSparseQR(pivoting=1, right_solver=1).compute_solve(mat, b), ref.solve(b)) <= TOL_X


def test_multi_gpu_exchange_needs_the_tsqr_path(qk):
    import ctypes as C
    from qrkit_b200 import capi
    d = capi.QrkDesc()
    d.kind, d.num_blocks, d.block_rows, d.block_cols, d.border_cols = capi.QRK_BLOCK_ANGULAR, 10, 2, 1, 9
    h = C.c_void_p()
    assert capi.lib().qrk_create(C.byref(d), C.byref(h)) == capi.QRK_STATUS_OK
    assert capi.lib().qrk_angular_set_world(h, 2) == capi.QRK_STATUS_INVALID_ARGUMENT     # wide borders: single GPU for now
    capi.lib().qrk_destroy(h)


def test_config3_full_size_properties(qk, oracle):
    """BASELINE config 3: ellipse Jacobian at N = 1M (2M x (1M+5)).  Size-independent checks: normal equations of the
    least-squares solution, x2 against a float128-free reference (numpy lstsq on the Schur-reduced 5-column problem)."""
    n = 1_000_000
    J1, J2, rhs = ellipse_problem(n)
    mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(J1, block_rows=2, block_cols=1), J2)
    s = qk.BlockAngularSparseQR(pivoting=0)
    x = s.compute_solve(mat, rhs)
    x1, x2 = x[:n], x[n:]
    # eliminate the diagonal blocks analytically: for each point, project out the 2x1 block direction
    a = J1.reshape(n, 2)
    nrm2 = (a * a).sum(axis=1)
    Jb = J2.reshape(n, 2, 5)
    proj = np.einsum("ni,nij->nj", a, Jb) / nrm2[:, None]
    Ared = (Jb - a[:, :, None] * proj[:, None, :]).reshape(2 * n, 5)
    rb = rhs.reshape(n, 2)
    rred = (rb - a * ((a * rb).sum(axis=1) / nrm2)[:, None]).reshape(-1)
    x2_ref = np.linalg.lstsq(Ared, rred, rcond=None)[0]
    kappa = np.linalg.cond(Ared)                                  # 16.1: the reduced problem is well conditioned
    assert kappa < 100
    assert rel(x2, x2_ref) <= TOL_X
    x1_ref = ((a * (rb - np.einsum("nij,j->ni", Jb, x2))).sum(axis=1)) / nrm2
    assert rel(x1, x1_ref) <= TOL_X
    # the whole problem against the oracle (the reference's algorithm: explicit sparse Q1, dense ColPiv QR of the 1M x 5
    # residual) at the north-star tolerance, P_c and rank included
    ref = oracle.BlockAngularOracle(J2, br=np.full(n, 2, dtype=np.int32), bc=np.full(n, 1, dtype=np.int32), values=J1, left_colpiv=False, right_kind=0)
    assert rel(x, ref.solve(rhs)) <= TOL_X
    assert np.array_equal(s.colsPermutation(), ref.colsPermutation()) and s.rank() == ref.rank
    # window against the oracle's packed left factor is covered by the block-diagonal tests; here the residual's
    # orthogonality to every column: A^T (A x - b) = 0
    res = (a * x1[:, None] + np.einsum("nij,j->ni", Jb, x2) - rb)
    g1 = (a * res).sum(axis=1)
    g2 = np.einsum("nij,ni->j", Jb, res)
    scale = np.abs(J2).max() * np.linalg.norm(rhs)
    assert np.abs(g1).max() <= 1e-9 * scale and np.abs(g2).max() <= 1e-7 * scale


@pytest.mark.parametrize("panel", [2, 3])
@pytest.mark.parametrize("nb,r,c,m2", [(40, 7, 2, 9), (64, 8, 4, 16), (200, 2, 1, 5)])
def test_thin_sparse_right_solver_vs_oracle(qk, oracle, nb, r, c, m2, panel):
    """RightSolver = BlockedThinSparseQR<.., SuggestedBlockCols = panel> (test/test-qrkit.cpp:54-57, 329-362): ColPiv inside
    every panel of the border, full column rank.  P_c bit-exact, R index arrays bit-exact, R2 up to row signs, x to 1e-10."""
    vals = uniform_blocks(nb, r, c)
    J2 = dense_border(nb * r, m2)
    b = vector(nb * r, seed=5)
    m1 = nb * c
    mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), J2)
    ref = oracle.BlockAngularOracle(J2, br=np.full(nb, r), bc=np.full(nb, c), values=vals, left_colpiv=True, right_kind=2, panel=panel)
    s = qk.BlockAngularSparseQR(mat, pivoting=1, right_solver=2, right_panel=panel)
    assert s.rank() == ref.rank == m1 + m2 and s.info() == qk.QRK_INFO_SUCCESS
    P = s.colsPermutation()
    assert np.array_equal(P, ref.colsPermutation()), "per-panel pivot order must be bit-exact"
    assert not np.array_equal(P[m1:], m1 + np.arange(m2)), "the case should exercise pivoting inside a panel"
    R, Rr = s.matrixR(), ref.matrixR()
    assert np.array_equal(R.outer, Rr.outer) and np.array_equal(R.inner, Rr.inner)
    Rd, Rrd = R.toarray(), Rr.toarray()
    assert rel(Rd[:m1, :], Rrd[:m1, :]) <= TOL_R
    assert rel(_sign_fix(Rd[m1:m1 + m2, m1:], Rrd[m1:m1 + m2, m1:]), Rrd[m1:m1 + m2, m1:]) <= TOL_R * 10
    x_ref = ref.solve(b)
    assert rel(s.solve(b), x_ref) <= TOL_X
    assert rel(qk.BlockAngularSparseQR(pivoting=1, right_solver=2, right_panel=panel).compute_solve(mat, b), x_ref) <= TOL_X
    A = np.hstack([blocks_to_dense(vals, np.full(nb, r), np.full(nb, c)), J2])
    assert rel(s.solve(b), np.linalg.lstsq(A, b, rcond=None)[0]) <= TOL_X


def test_thin_sparse_right_solver_rank_deficient_border(qk, oracle):
    """BlockedThinSparseQR's zero-pivot deferral (BlockedThinSparseQR.h:250-255, 150-158): border columns that are exactly zero
    are detected inside their panel, moved to the end of P2 and excluded from the rank; x is the basic solution (zero on the
    deferred columns).  Deferral order, P_c and rank bit-exact against the oracle; A P = Q R through R^T R = (A P)^T (A P)."""
    nb, r, c, m2, panel = 50, 7, 2, 11, 2
    vals = uniform_blocks(nb, r, c)
    J2 = dense_border(nb * r, m2)
    J2[:, 2] = 0.0; J2[:, 5] = 0.0; J2[:, 7] = 0.0           # first / second column of their panels (a panel that is ALL zero
                                                             # is not flagged by Eigen's rule: its threshold is 0 and 0 < 0 fails)
    b = vector(nb * r, seed=5)
    m1 = nb * c
    mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), J2)
    ref = oracle.BlockAngularOracle(J2, br=np.full(nb, r), bc=np.full(nb, c), values=vals, left_colpiv=True, right_kind=2, panel=panel)
    s = qk.BlockAngularSparseQR(pivoting=1, right_solver=2, right_panel=panel)
    x = s.compute_solve(mat, b)
    assert s.rank() == ref.rank == m1 + m2 - 3
    P = s.colsPermutation()
    assert np.array_equal(P, ref.colsPermutation())
    assert list(P[-3:] - m1) == [2, 5, 7]                     # deferral order = panel order
    Rd = s.matrixR().toarray()[:m1 + m2, :]
    A = np.hstack([blocks_to_dense(vals, np.full(nb, r), np.full(nb, c)), J2])
    AP = A[:, P]
    assert rel(Rd.T @ Rd, AP.T @ AP) <= 1e-12
    assert np.all(x[m1 + np.array([2, 5, 7])] == 0.0)
    keep = np.setdiff1d(np.arange(m1 + m2), m1 + np.array([2, 5, 7]))
    assert rel(x[keep], np.linalg.lstsq(A[:, keep], b, rcond=None)[0]) <= TOL_X
    assert rel(x, ref.solve(b)) <= TOL_X
    with pytest.raises(qk.QrkError):                            # stored-factor products are refused once columns were deferred
        s.solve(b)


def test_fused_peer_exchange_timeout_poisons_and_reports(qk, oracle):
    """A peer that never delivers its triangle: the waiting root kernel gives up after qrk_angular_p2p_set_timeout seconds,
    x comes back NaN (never a plausible-looking partial solve), the synchronous call returns QRK_STATUS_PEER_TIMEOUT and
    qrk_angular_p2p_status reports it; a fresh attach clears the condition and the same handles then solve correctly."""
    import ctypes as C
    import torch
    from qrkit_b200 import capi
    from qrkit_b200.capi import QRK_DEVICE, QRK_HOST, QrkDesc, check
    L = capi.lib()
    n, world = 2000, 2
    per = n // world
    J1, J2, rhs = ellipse_problem(n)
    ref = oracle.BlockAngularOracle(J2, br=np.full(n, 2), bc=np.full(n, 1), values=J1, left_colpiv=False, right_kind=0)
    x_ref = ref.solve(rhs)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    vp = lambda t: C.c_void_p(t.data_ptr())
    hs, bufs, xbufs = [], [], []
    for g in range(world):
        rows = slice(2 * g * per, 2 * (g + 1) * per)
        dJ1, dJ2 = dev(J1[rows]), dev(J2[rows, :].T)
        d = QrkDesc()
        d.kind, d.num_blocks, d.block_rows, d.block_cols, d.pivoting, d.border_cols = capi.QRK_BLOCK_ANGULAR, per, 2, 1, 0, 5
        h = C.c_void_p()
        check(L.qrk_create(C.byref(d), C.byref(h)))
        check(L.qrk_angular_set_world(h, world), h)
        check(L.qrk_set_border(h, vp(dJ2), 2 * per, QRK_DEVICE), h)
        check(L.qrk_set_blocks(h, vp(dJ1), QRK_DEVICE), h)
        xb, nb_ = C.c_void_p(), C.c_int64()
        check(L.qrk_angular_xchg_buffer(h, C.byref(xb), C.byref(nb_)), h)
        hs.append(h); bufs.append((dJ1, dJ2)); xbufs.append(xb.value)
    peers = (C.c_void_p * world)(*xbufs)
    for g, h in enumerate(hs):
        check(L.qrk_angular_p2p_attach(h, peers, world, g), h)
    assert L.qrk_angular_p2p_set_timeout(hs[0], C.c_double(0.0)) == capi.QRK_STATUS_INVALID_ARGUMENT
    check(L.qrk_angular_p2p_set_timeout(hs[0], C.c_double(0.05)), hs[0])

    # rank 1 never runs: rank 0 must come back within the timeout, not hang
    hb = np.ascontiguousarray(rhs[:2 * per]); hx = np.zeros(per + 5); hv = np.ascontiguousarray(J1[:2 * per])
    hp = lambda a: a.ctypes.data_as(C.c_void_p)
    t0 = time.perf_counter()
    st = L.qrk_compute_solve(hs[0], hp(hv), hp(hb), hp(hx), QRK_HOST)
    assert time.perf_counter() - t0 < 5.0
    assert st == capi.QRK_STATUS_PEER_TIMEOUT, (st, L.qrk_last_error(hs[0]))
    assert np.isnan(hx[per:]).all()
    to = C.c_int32(0)
    check(L.qrk_angular_p2p_status(hs[0], C.byref(to)), hs[0])
    assert to.value != 0

    # recovery: re-attach both ranks (resets flags, parities and the error word), then a normal two-rank step
    for g, h in enumerate(hs):
        check(L.qrk_angular_p2p_attach(h, peers, world, g), h)
    dbs = [dev(rhs[2 * g * per:2 * (g + 1) * per]) for g in range(world)]
    dxs = [torch.zeros(per + 5, dtype=torch.float64, device="cuda") for _ in range(world)]
    for g, h in enumerate(hs):
        check(L.qrk_compute_solve(h, vp(bufs[g][0]), vp(dbs[g]), vp(dxs[g]), QRK_DEVICE), h)
    for g, h in enumerate(hs):
        check(L.qrk_angular_p2p_status(h, C.byref(to)), h)
        assert to.value == 0
        xg = dxs[g].cpu().numpy()
        assert rel(xg[:per], x_ref[g * per:(g + 1) * per]) <= 1e-9
        assert rel(xg[per:], x_ref[n:]) <= 1e-9
    for h in hs:
        L.qrk_destroy(h)


@pytest.mark.parametrize("m2,nb", [(65, 40), (100, 60), (130, 70), (383, 128)])
def test_register_resident_colpiv_triangle_vs_oracle(qk, oracle, m2, nb):
    """Borders of 65..384 columns with the ColPiv right solver take the register-resident cluster kernel for the pivoted QR of
    the border's triangle (dense_tri_reg.cuh: columns never move, pivoting exchanges logical positions): ragged widths (not a
    multiple of 8, 32 or 64 columns; 383 = one short of the limit), the full check of _check (P2 bit-exact, R, x, Q identities)."""
    r, c = 7, 2
    vals = uniform_blocks(nb, r, c)
    J2 = dense_border(nb * r, m2, seed=SEED_A + 70 + m2)
    b = vector(nb * r, seed=5)
    _check(qk, oracle, vals, r, c, J2, b, piv=1)


def test_register_resident_colpiv_triangle_rank_deficient(qk, oracle):
    """The same kernel on a rank-deficient border with exact zero columns: Eigen's nonzero-pivot threshold and the first-maximum
    rule decide rank and P2 exactly as in the oracle's restatement, the zero columns end up behind the rank in index order.
    (Duplicated or linearly dependent columns are not used here: the pivoted QR runs on the TRIANGLE of the unpivoted first
    stage, where their norms tie -- and their residuals vanish -- only up to rounding, so neither their order nor whether a
    1e-13 residual counts as a pivot is defined bit for bit; the reference decides both on the tall matrix.)"""
    r, c, nb, m2 = 7, 2, 40, 96
    vals = uniform_blocks(nb, r, c)
    J2 = dense_border(nb * r, m2, seed=SEED_A + 91)
    for j in (7, 50, 51, 95):
        J2[:, j] = 0.0
    b = vector(nb * r, seed=9)
    mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), J2)
    ref = oracle.BlockAngularOracle(J2, br=np.full(nb, r), bc=np.full(nb, c), values=vals, left_colpiv=True, right_kind=0)
    s = qk.BlockAngularSparseQR(mat, pivoting=1)
    rank = nb * c + m2 - 4
    assert s.rank() == ref.rank == rank
    assert np.array_equal(s.colsPermutation(), ref.colsPermutation())
    assert rel(s.solve(b), ref.solve(b)) <= 1e-9
    assert rel(qk.BlockAngularSparseQR(pivoting=1).compute_solve(mat, b), ref.solve(b)) <= 1e-9


@pytest.mark.parametrize("m2,right", [(24, 0), (96, 0), (40, 1)])
def test_wide_border_step_replayed_from_cuda_graph(qk, oracle, m2, right):
    """The wide-border step is captured into a CUDA graph the second time a handle sees the same buffers and replayed from
    then on (capi.cu: wide_run).  Four successive compute_solve calls on ONE handle with different matrices and right-hand
    sides (host memspace: the handle's staging buffers are the stable device pointers): every result is bit-identical to a
    fresh handle's eager first call on the same data, and the first one matches the oracle."""
    r, c, nb = 7, 2, 60
    s = qk.BlockAngularSparseQR(pivoting=1, right_solver=right)
    for it in range(4):
        vals = uniform_blocks(nb, r, c, seed=SEED_A + 3 * it)
        J2 = dense_border(nb * r, m2, seed=SEED_A + 500 + it)
        b = vector(nb * r, seed=40 + it)
        mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), J2)
        x = s.compute_solve(mat, b).copy()
        x_fresh = qk.BlockAngularSparseQR(pivoting=1, right_solver=right).compute_solve(mat, b)
        assert np.array_equal(x, x_fresh), f"call {it}: graph replay differs from the eager step"
        assert s.rank() == nb * c + m2
        if it == 0:
            ref = oracle.BlockAngularOracle(J2, br=np.full(nb, r), bc=np.full(nb, c), values=vals, left_colpiv=True,
                                            right_kind=right, **({"panel": 2} if right == 1 else {}))
            assert rel(x, ref.solve(b)) <= TOL_X


@pytest.mark.parametrize("piv", [0, 1])
def test_tsqr_step_replayed_from_cuda_graph(qk, oracle, piv):
    """The narrow-border step (K1 -> TSQR root -> K3, the last a programmatic dependent launch) is captured into a CUDA graph
    the second time a handle sees the same buffers and replayed from then on: five successive compute_solve calls on ONE handle
    with different Jacobians and right-hand sides are bit-identical to a fresh handle's eager call, rank and P_c included, and
    compute() + solve() still works on the handle afterwards (host-side state restored by the replay)."""
    n = 3001
    s = qk.BlockAngularSparseQR(pivoting=piv)
    for it in range(5):
        J1, J2, rhs = ellipse_problem(n)
        J1 = J1 * (1.0 + 0.01 * it); J2 = J2 + 0.001 * it * dense_border(2 * n, 5, seed=SEED_A + it); rhs = rhs + 0.1 * it
        mat = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(J1, block_rows=2, block_cols=1), J2)
        x = s.compute_solve(mat, rhs).copy()
        fresh = qk.BlockAngularSparseQR(pivoting=piv)
        assert np.array_equal(x, fresh.compute_solve(mat, rhs)), f"call {it}: graph replay differs from the eager step"
        assert s.rank() == fresh.rank() == n + 5
        assert np.array_equal(s.colsPermutation(), fresh.colsPermutation())
    ref = oracle.BlockAngularOracle(J2, br=np.full(n, 2), bc=np.full(n, 1), values=J1, left_colpiv=bool(piv), right_kind=0)
    assert rel(x, ref.solve(rhs)) <= TOL_X
    s.compute(mat)
    assert rel(s.solve(rhs), ref.solve(rhs)) <= TOL_X
    # the reference's own calling pattern, compute(J) then solve(b), four rounds on the same handle: both calls replay
    for it in range(4):
        J1b = J1 * (1.0 + 0.02 * it); rhsb = rhs - 0.05 * it
        matb = qk.BlockMatrix1x2(qk.SparseBlockDiagonal(J1b, block_rows=2, block_cols=1), J2)
        s.compute(matb)
        xs = s.solve(rhsb).copy()
        fresh = qk.BlockAngularSparseQR(matb, pivoting=piv)
        assert np.array_equal(xs, fresh.solve(rhsb)), f"round {it}: replayed compute() + solve() differs from the eager pair"
        assert np.array_equal(s.solve(rhsb), xs)                     # a second solve on the same factorisation
