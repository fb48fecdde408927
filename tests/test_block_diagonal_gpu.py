"""GPU parity tests of the block-diagonal hot path (run with -m gpu on the B200 box).

Every call goes through the C ABI (libqrkit_b200.so); the CPU oracle (oracle/) is the checker only.
Tolerances are the north-star ones (BASELINE.json): R within 1e-12 relative Frobenius, ||QR-A||/||A|| <= 1e-13,
x within 1e-10 relative; permutations and sparse index arrays bit-exact.
Shapes follow the reference's tests (test/test-qrkit.cpp:167-206: 256 blocks of 7x2, ColPiv) and BASELINE configs 2 and 5."""
import numpy as np
import pytest

from helpers import SEED_A, blocks_to_dense, rel, splitmix64, synth, uniform_blocks, vector

pytestmark = pytest.mark.gpu

TOL_R, TOL_QR, TOL_X = 1e-12, 1e-13, 1e-10

SMALL_SHAPES = [(2, 1), (3, 1), (4, 2), (6, 3), (7, 2), (8, 2), (8, 4), (9, 2)]          # thread-per-block kernels
GENERIC_SHAPES = [(5, 3), (4, 4), (1, 1), (16, 8), (33, 7), (32, 16), (64, 32), (128, 64), (60, 50)]   # team-per-block
# unpivoted blocks wider than one 8-column panel take the blocked compact-WY / DMMA kernel (bd_wy.cuh): every row-per-lane
# class (rp <= 32 / 64 / 128), ragged rows and columns (zero padding), square blocks, one- and two-warp tile schedules
WY_SHAPES = [(48, 24), (80, 40), (96, 48), (112, 56), (20, 12), (40, 9), (127, 63), (100, 100), (17, 16), (24, 24),
             (128, 16), (31, 30), (72, 64), (128, 128)]


@pytest.fixture(scope="module")
def qk():
    import qrkit_b200 as q
    if q.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests must run on the B200 box (there is no CPU fallback)")
    return q


def _batched_q(packed, tau, nb, r, c):
    """Explicit Q_i (nb, r, r) from packed reflectors, numpy restatement of H_0...H_{c-1} (checker side)."""
    P = packed.reshape(nb, c, r).transpose(0, 2, 1)       # (nb, r, c)
    T = tau.reshape(nb, c)
    Q = np.broadcast_to(np.eye(r), (nb, r, r)).copy()
    for k in range(min(r, c) - 1, -1, -1):
        v = np.zeros((nb, r))
        v[:, k] = 1.0
        v[:, k + 1:] = P[:, k + 1:, k]
        w = np.einsum("bi,bij->bj", v, Q)
        Q -= T[:, k, None, None] * v[:, :, None] * w[:, None, :]
    return Q


def _check_uniform(qk, orc, nb, r, c, piv, lo=0.5, hi=5.0, seed=SEED_A):
    vals = uniform_blocks(nb, r, c, seed=seed, lo=lo, hi=hi)
    b = vector(nb * r, seed=seed + 5)
    mat = qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c)
    solver = qk.BlockDiagonalSparseQR(pivoting=piv)
    x_fused = solver.compute_solve(mat, b)
    ref = orc.bd_compact_uniform(nb, r, c, vals, b, colpiv=bool(piv))
    pk, tau = solver.packed()
    perm = solver.colsPermutation()
    assert np.array_equal(perm, ref["perm"]), "column permutation must be bit-exact"
    assert rel(pk, ref["packed"]) <= TOL_R
    assert rel(tau, ref["tau"]) <= TOL_R
    assert rel(x_fused, ref["x"]) <= TOL_X
    # ||Q R - A P|| / ||A|| per batch
    Q = _batched_q(pk, tau, nb, r, c)
    Rm = np.triu(pk.reshape(nb, c, r).transpose(0, 2, 1))
    A = vals.reshape(nb, c, r).transpose(0, 2, 1)
    local = perm.reshape(nb, c) - (np.arange(nb) * c)[:, None]
    AP = np.take_along_axis(A, local[:, None, :].repeat(r, axis=1), axis=2)
    assert rel(np.einsum("bij,bjk->bik", Q, Rm), AP) <= TOL_QR
    # two-call path (compute, then solve) agrees with the fused path
    solver2 = qk.BlockDiagonalSparseQR(mat, pivoting=piv)
    x2 = solver2.solve(b)
    assert rel(x2, ref["x"]) <= TOL_X
    assert solver2.rank() == nb * c and solver2.info() == qk.QRK_INFO_SUCCESS
    return solver2, vals, b, ref


@pytest.mark.parametrize("r,c", SMALL_SHAPES)
@pytest.mark.parametrize("piv", [0, 1])
def test_small_blocks_vs_oracle(qk, oracle, r, c, piv):
    for nb in (1, 127, 128, 129, 1000):
        _check_uniform(qk, oracle, nb, r, c, piv)


@pytest.mark.parametrize("r,c", [(8, 4), (7, 2)])
@pytest.mark.parametrize("piv", [0, 1])
def test_small_blocks_signed_inputs(qk, oracle, r, c, piv):
    _check_uniform(qk, oracle, 513, r, c, piv, lo=-1.0, hi=1.0, seed=SEED_A + 77)


@pytest.mark.parametrize("r,c", GENERIC_SHAPES)
@pytest.mark.parametrize("piv", [0, 1])
def test_generic_uniform_blocks_vs_oracle(qk, oracle, r, c, piv):
    nb = 37 if r * c > 1000 else 301
    _check_uniform(qk, oracle, nb, r, c, piv)
    _check_uniform(qk, oracle, 3, r, c, piv, lo=-1.0, hi=1.0, seed=SEED_A + 3)


@pytest.mark.parametrize("r,c", WY_SHAPES)
def test_blocked_wy_blocks_vs_oracle(qk, oracle, r, c):
    _check_uniform(qk, oracle, 41, r, c, 0)
    _check_uniform(qk, oracle, 3, r, c, 0, lo=-1.0, hi=1.0, seed=SEED_A + 11)


def _mixed_problem(nb, seed=SEED_A):
    """BASELINE config 5 generator: r = 32 + 16*(hash(i) mod 7), c = r/2."""
    from helpers import splitmix64
    hsh = splitmix64(np.arange(nb, dtype=np.uint64) ^ np.uint64(seed))
    br = (32 + 16 * (hsh % np.uint64(7)).astype(np.int64)).astype(np.int32)
    bc = (br // 2).astype(np.int32)
    vals = np.concatenate([uniform_blocks(1, int(r), int(c), seed=seed, block0=i) for i, (r, c) in enumerate(zip(br, bc))])
    return br, bc, vals


@pytest.mark.parametrize("piv", [0, 1])
def test_mixed_blocks_vs_oracle(qk, oracle, piv):
    nb = 60
    br, bc, vals = _mixed_problem(nb)
    b = vector(int(br.sum()), seed=17)
    mat = qk.SparseBlockDiagonal(vals, rows=br, cols=bc)
    solver = qk.BlockDiagonalSparseQR(pivoting=piv)
    x_fused = solver.compute_solve(mat, b)
    ref = oracle.BlockDiagonalOracle(br, bc, vals, colpiv=bool(piv))
    pk_ref, tau_ref = ref.packed()
    pk, tau = solver.packed()
    assert np.array_equal(solver.colsPermutation(), ref.colsPermutation())
    assert rel(pk, pk_ref) <= TOL_R and rel(tau, tau_ref) <= TOL_R
    x_ref = ref.solve(b)
    assert rel(x_fused, x_ref) <= TOL_X
    assert rel(solver.solve(b), x_ref) <= TOL_X
    # sparse R: index arrays bit-exact, values to tolerance
    R, Rr = solver.matrixR(), ref.matrixR()
    assert np.array_equal(R.outer, Rr.outer) and np.array_equal(R.inner, Rr.inner)
    assert rel(R.values, Rr.val) <= TOL_R
    # Q^T b in the FullQ layout
    assert rel(solver.applyQt(b), ref.apply_qt(b)) <= TOL_R
    A = blocks_to_dense(vals, br, bc)
    P = solver.colsPermutation()
    y = solver.applyQ(np.asfortranarray(R.toarray()))
    assert rel(y, A[:, P]) <= TOL_QR
    # the thin factor Q1 = (A P) R^-1: Q1^T b = (Q^T b)[0:n_cols], Q1 (R x) = A P x
    n_cols = int(bc.sum())
    assert np.array_equal(solver.applyQtThin(b), solver.applyQt(b)[:n_cols])
    xs = vector(n_cols, seed=23)
    assert rel(solver.applyQThin(R.toarray()[:n_cols, :] @ xs), A[:, P] @ xs) <= TOL_QR


@pytest.mark.parametrize("qformat", [0, 1])
def test_reference_test0_properties(qk, oracle, qformat):
    """test/test-qrkit.cpp:167-206: 256 blocks of 7x2 (1792 x 512), ColPiv per block.
    Q*R = A*P, Q^T*(A*P) = R, x recovered — plus bit-exact Q/R index arrays against the oracle."""
    nb, r, c = 256, 7, 2
    vals = uniform_blocks(nb, r, c)
    mat = qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c)
    solver = qk.BlockDiagonalSparseQR(mat, pivoting=qk.QRK_PIVOT_COLPIV, q_format=qformat)
    ref = oracle.BlockDiagonalOracle(np.full(nb, r), np.full(nb, c), vals, colpiv=True, qformat=qformat)
    assert solver.info() == qk.QRK_INFO_SUCCESS and solver.rank() == ref.rank
    assert solver.rows() == 1792 and solver.cols() == 512
    Q, R = solver.matrixQ(), solver.matrixR()
    Qr, Rr = ref.matrixQ(), ref.matrixR()
    assert np.array_equal(Q.outer, Qr.outer) and np.array_equal(Q.inner, Qr.inner)
    assert np.array_equal(R.outer, Rr.outer) and np.array_equal(R.inner, Rr.inner)
    assert rel(Q.values, Qr.val) <= TOL_R and rel(R.values, Rr.val) <= TOL_R
    P = solver.colsPermutation()
    assert np.array_equal(P, ref.colsPermutation())
    assert np.array_equal(solver.rowsPermutation(), np.arange(1792, dtype=np.int32))
    A = blocks_to_dense(vals, np.full(nb, r), np.full(nb, c))
    Qd, Rd = Q.toarray(), R.toarray()
    assert rel(Qd @ Rd, A[:, P]) <= TOL_QR
    assert rel(Qd.T @ A[:, P], Rd) <= TOL_QR
    assert rel(Qd.T @ Qd, np.eye(1792)) <= TOL_QR
    # operator forms of matrixQ().transpose()*B and matrixQ()*B on several right-hand sides
    B = synth(SEED_A + 9, 3, np.arange(1792)[:, None], np.arange(3)[None, :], -1.0, 1.0)
    assert rel(solver.applyQt(B), Qd.T @ B) <= TOL_R
    assert rel(solver.applyQ(B), Qd @ B) <= TOL_R
    if qformat == 0:   # solve is only meaningful for FullQ (R upper triangular), BlockDiagonalSparseQR.h:134-135
        x_true = vector(512, seed=21)
        b = A @ x_true
        x = solver.solve(b)
        assert rel(x, x_true) <= TOL_X
        assert rel(x, ref.solve(b)) <= TOL_X
        X = solver.solve(np.stack([b, 2 * b], axis=1))
        assert rel(X[:, 1], 2 * x_true) <= TOL_X


def test_general_sparse_matrix_through_structure_detection(qk):
    """SparseBlockDiagonal::fromSparseMatrix (SparseBlockDiagonal.h:96-130) in front of the solver: a row-shuffled
    block-diagonal sparse matrix is reordered (AsBandedAsPossible), its blocks detected and extracted on the host, then
    factored on the GPU; the caller applies rowsPermutation() to b as in test/test-qrkit.cpp:235."""
    import scipy.sparse as sp
    from qrkit_b200 import structure
    nb, r, c = 300, 7, 2
    vals0 = uniform_blocks(nb, r, c)
    A0 = sp.csr_matrix(blocks_to_dense(vals0, np.full(nb, r), np.full(nb, c)))
    shuffle = np.random.default_rng(3).permutation(nb * r)
    A = A0[shuffle, :]
    vals, br, bc, blocks, perm, has = structure.from_sparse_matrix(A)
    assert has and len(br) == nb and np.all(br == r) and np.all(bc == c)
    x_true = vector(nb * c, seed=31)
    b = A @ x_true
    pb = np.empty_like(b)
    pb[perm] = b                                          # (P b)[perm[i]] = b[i]
    solver = qk.BlockDiagonalSparseQR(qk.SparseBlockDiagonal(vals, rows=br, cols=bc), pivoting=1)
    assert rel(solver.solve(pb), x_true) <= TOL_X
    b2 = vector(nb * r, seed=32)
    pb2 = np.empty_like(b2); pb2[perm] = b2
    assert rel(solver.solve(pb2), np.linalg.lstsq(A.toarray(), b2, rcond=None)[0]) <= TOL_X


def test_zero_block_tail_rows(qk, oracle):
    """Rows below the last block get Q(i,i)=1 (BlockDiagonalSparseQR.h:530-533)."""
    nb, r, c = 10, 7, 2
    vals = uniform_blocks(nb, r, c)
    n_rows = nb * r + 5
    mat = qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c, n_rows=n_rows)
    solver = qk.BlockDiagonalSparseQR(mat, pivoting=qk.QRK_PIVOT_COLPIV)
    ref = oracle.BlockDiagonalOracle(np.full(nb, r), np.full(nb, c), vals, n_rows=n_rows, colpiv=True)
    Q, Qr = solver.matrixQ(), ref.matrixQ()
    assert np.array_equal(Q.outer, Qr.outer) and np.array_equal(Q.inner, Qr.inner) and rel(Q.values, Qr.val) <= TOL_R
    b = vector(n_rows, seed=3)
    assert rel(solver.applyQt(b), ref.apply_qt(b)) <= TOL_R
    assert rel(solver.solve(b), ref.solve(b)) <= TOL_X


def test_landscape_blocks_are_invalid_input(qk):
    """BlockDiagonalSparseQR.h:509-516: blocks with more columns than rows -> info() == InvalidInput."""
    vals = uniform_blocks(4, 2, 3)
    solver = qk.BlockDiagonalSparseQR(qk.SparseBlockDiagonal(vals, block_rows=2, block_cols=3))
    assert solver.info() == qk.QRK_INFO_INVALID_INPUT


def test_not_factorized_is_reported(qk):
    import ctypes as C
    from qrkit_b200 import capi
    d = capi.QrkDesc()
    d.kind, d.num_blocks, d.block_rows, d.block_cols = 0, 4, 8, 4
    h = C.c_void_p()
    capi.check(capi.lib().qrk_create(C.byref(d), C.byref(h)))
    v = C.c_int64()
    assert capi.lib().qrk_rank(h, C.byref(v)) == capi.QRK_STATUS_NOT_FACTORIZED
    x = np.zeros(16)
    assert capi.lib().qrk_solve(h, x.ctypes.data_as(C.c_void_p), 32, x.ctypes.data_as(C.c_void_p), 16, 1, 0) == capi.QRK_STATUS_NOT_FACTORIZED
    capi.lib().qrk_destroy(h)


def test_degenerate_columns(qk, oracle):
    """Zero tails (tau = 0, no sign flip), exact ties between column norms (first maximum wins)."""
    r, c = 8, 4
    blocks = []
    a = np.zeros((r, c)); a[:4, :4] = np.triu(np.arange(1.0, 17.0).reshape(4, 4)) + np.eye(4)   # already triangular
    blocks.append(a)
    t = synth(5, 0, np.arange(r)[:, None], np.arange(c)[None, :])
    # columns 0 and 1 have bit-identical norms (same squares in the same order) but are independent
    t[:, 1] = t[:, 0] * np.array([1, -1, 1, -1, 1, 1, -1, -1.0])
    t[:, 2] *= 0.1; t[:, 3] *= 0.01                                                        # clearly separated afterwards
    blocks.append(t)
    blocks.append(-synth(6, 0, np.arange(r)[:, None], np.arange(c)[None, :]))
    vals = np.concatenate([blk.T.reshape(-1) for blk in blocks])
    nb = len(blocks)
    b = vector(nb * r, seed=4)
    for piv in (0, 1):
        solver = qk.BlockDiagonalSparseQR(pivoting=piv)
        x = solver.compute_solve(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), b)
        ref = oracle.bd_compact_uniform(nb, r, c, vals, b, colpiv=bool(piv))
        pk, tau = solver.packed()
        assert np.array_equal(solver.colsPermutation(), ref["perm"])
        assert rel(pk, ref["packed"]) <= TOL_R and rel(tau, ref["tau"]) <= TOL_R
        assert rel(x, ref["x"]) <= TOL_X
        if piv == 0:
            assert tau[0] == 0.0 and pk[0] == a[0, 0]


def test_empty_problem(qk):
    solver = qk.BlockDiagonalSparseQR(qk.SparseBlockDiagonal(np.zeros(0), num_blocks=0, block_rows=8, block_cols=4))
    assert solver.rows() == 0 and solver.cols() == 0 and solver.rank() == 0
    assert solver.solve(np.zeros(0)).shape == (0,)


@pytest.mark.parametrize("piv", [0, 1])
def test_headline_config_full_size_properties(qk, oracle, piv):
    """BASELINE config 2 at full size (1M blocks of 8x4): size-independent properties over all blocks
    (x recovered from a consistent system, ||QR - AP||/||A||, orthogonality) and the oracle on a 20k-block window."""
    nb, r, c = 1_000_000, 8, 4
    vals = uniform_blocks(nb, r, c)
    x_true = vector(nb * c, seed=SEED_A + 1)
    A = vals.reshape(nb, c, r).transpose(0, 2, 1)
    b = np.einsum("bij,bj->bi", A, x_true.reshape(nb, c)).reshape(-1)
    solver = qk.BlockDiagonalSparseQR(pivoting=piv)
    x = solver.compute_solve(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), b)
    assert rel(x, x_true) <= TOL_X
    pk, tau = solver.packed()
    perm = solver.colsPermutation()
    local = perm.reshape(nb, c) - (np.arange(nb) * c)[:, None]
    assert np.array_equal(np.sort(local, axis=1), np.broadcast_to(np.arange(c), (nb, c)))     # a permutation per block
    Q = _batched_q(pk, tau, nb, r, c)
    Rm = np.triu(pk.reshape(nb, c, r).transpose(0, 2, 1))
    AP = np.take_along_axis(A, local[:, None, :].repeat(r, axis=1), axis=2)
    assert rel(np.einsum("bij,bjk->bik", Q, Rm), AP) <= TOL_QR
    w0 = 490_000
    win = slice(w0 * r * c, (w0 + 20_000) * r * c)
    ref = oracle.bd_compact_uniform(20_000, r, c, vals[win], b[w0 * r:(w0 + 20_000) * r], colpiv=bool(piv))
    assert rel(pk[win], ref["packed"]) <= TOL_R
    assert np.array_equal(local[w0:w0 + 20_000].reshape(-1), ref["perm"] - np.repeat(np.arange(20_000) * c, c))
    assert rel(x[w0 * c:(w0 + 20_000) * c], ref["x"]) <= TOL_X


def test_config5_full_size_properties(qk):
    """BASELINE config 5 at full size: 100k mixed blocks 32x16 .. 128x64 (2.97 GB of values, generated on the device).
    Size-independent checks through the C ABI with device pointers: x recovered from a consistent system, and the normal
    equations A^T (A x - b) = 0 of a least-squares right-hand side, per size class with batched products."""
    import ctypes as C
    import torch
    from qrkit_b200 import capi
    from qrkit_b200.capi import QRK_DEVICE, QrkDesc, check
    L = capi.lib()
    nb = 100_000
    hsh = splitmix64(np.arange(nb, dtype=np.uint64) ^ np.uint64(SEED_A))
    br = (32 + 16 * (hsh % np.uint64(7)).astype(np.int64)).astype(np.int32)
    bc = (br // 2).astype(np.int32)
    sizes = br.astype(np.int64) * bc
    voff = np.concatenate([[0], np.cumsum(sizes)])
    roff = np.concatenate([[0], np.cumsum(br.astype(np.int64))])
    coff = np.concatenate([[0], np.cumsum(bc.astype(np.int64))])
    total, rows, cols = int(voff[-1]), int(roff[-1]), int(coff[-1])
    vp = lambda t: C.c_void_p(t.data_ptr())
    A = torch.empty(total, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(vp(A), SEED_A, 0, total, 1, 0, 0.5, 5.0, None))
    x_true = torch.empty(cols, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(vp(x_true), SEED_A + 1, 0, cols, 1, 0, -1.0, 1.0, None))
    b_ls = torch.empty(rows, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(vp(b_ls), SEED_A + 5, 0, rows, 1, 0, -1.0, 1.0, None))
    torch.cuda.synchronize()

    def per_class():
        for r in sorted(set(br.tolist())):
            c = r // 2
            ids = np.nonzero(br == r)[0]
            vo = torch.from_numpy(voff[ids]).cuda(); ro = torch.from_numpy(roff[ids]).cuda(); co = torch.from_numpy(coff[ids]).cuda()
            blocks = A[vo[:, None] + torch.arange(r * c, device="cuda")[None, :]].reshape(len(ids), c, r).transpose(1, 2)   # nblk x r x c
            yield blocks, ro[:, None] + torch.arange(r, device="cuda")[None, :], co[:, None] + torch.arange(c, device="cuda")[None, :]

    def matvec(x):                                   # A x
        out = torch.empty(rows, dtype=torch.float64, device="cuda")
        for blocks, ri, ci in per_class():
            out[ri] = torch.bmm(blocks, x[ci].unsqueeze(2)).squeeze(2)
        return out

    def rmatvec(v):                                  # A^T v
        out = torch.empty(cols, dtype=torch.float64, device="cuda")
        for blocks, ri, ci in per_class():
            out[ci] = torch.bmm(blocks.transpose(1, 2), v[ri].unsqueeze(2)).squeeze(2)
        return out

    d = QrkDesc()
    d.kind, d.num_blocks, d.pivoting = 0, nb, 0
    d.rows = br.ctypes.data_as(C.POINTER(C.c_int32)); d.cols = bc.ctypes.data_as(C.POINTER(C.c_int32))
    h = C.c_void_p()
    check(L.qrk_create(C.byref(d), C.byref(h)))
    x = torch.empty(cols, dtype=torch.float64, device="cuda")
    b = matvec(x_true)
    torch.cuda.synchronize()
    check(L.qrk_compute_solve(h, vp(A), vp(b), vp(x), QRK_DEVICE), h)
    check(L.qrk_synchronize(h), h)
    assert float(torch.linalg.norm(x - x_true) / torch.linalg.norm(x_true)) <= 1e-10
    check(L.qrk_compute_solve(h, vp(A), vp(b_ls), vp(x), QRK_DEVICE), h)
    check(L.qrk_synchronize(h), h)
    g = rmatvec(matvec(x) - b_ls)
    assert float(g.abs().max() / rmatvec(b_ls).abs().max()) <= 1e-11
    L.qrk_destroy(h)
    # the oracle on windows of the full-size problem, unpivoted (blocked-WY / DMMA kernel) AND ColPiv (team-per-block kernel):
    # packed factors and tau at 1e-12, column permutations bit-exact, x at 1e-10 — 3 windows of 100 consecutive blocks each
    from oracle import oracle as orc
    packed = torch.empty(total, dtype=torch.float64, device="cuda")
    tau = torch.empty(cols, dtype=torch.float64, device="cuda")
    perm = torch.empty(cols, dtype=torch.int32, device="cuda")
    for piv in (0, 1):
        d = QrkDesc()
        d.kind, d.num_blocks, d.pivoting = 0, nb, piv
        d.rows = br.ctypes.data_as(C.POINTER(C.c_int32)); d.cols = bc.ctypes.data_as(C.POINTER(C.c_int32))
        h = C.c_void_p()
        check(L.qrk_create(C.byref(d), C.byref(h)))
        check(L.qrk_compute_solve(h, vp(A), vp(b_ls), vp(x), QRK_DEVICE), h)
        check(L.qrk_packed_factors(h, vp(packed), vp(tau), QRK_DEVICE), h)
        check(L.qrk_cols_permutation(h, vp(perm), QRK_DEVICE), h)
        check(L.qrk_synchronize(h), h)
        for w0 in (0, 49_950, nb - 100):
            ids = np.arange(w0, w0 + 100)
            v0, v1, r0, r1, c0, c1 = int(voff[w0]), int(voff[w0 + 100]), int(roff[w0]), int(roff[w0 + 100]), int(coff[w0]), int(coff[w0 + 100])
            ref = orc.BlockDiagonalOracle(br[ids], bc[ids], A[v0:v1].cpu().numpy(), colpiv=bool(piv))
            pk_ref, tau_ref = ref.packed()
            assert rel(packed[v0:v1].cpu().numpy(), pk_ref) <= TOL_R and rel(tau[c0:c1].cpu().numpy(), tau_ref) <= TOL_R
            assert np.array_equal(perm[c0:c1].cpu().numpy() - c0, ref.colsPermutation())
            assert rel(x[c0:c1].cpu().numpy(), ref.solve(b_ls[r0:r1].cpu().numpy())) <= TOL_X
        L.qrk_destroy(h)


def test_host_alloc_is_pinned_and_feeds_the_host_path(qk, oracle):
    """qrk_host_alloc hands out page-locked memory (cudaHostGetFlags succeeds on it) that the host-memspace one-call path
    accepts like any host pointer; the result equals the run from pageable numpy arrays bit for bit.  qrk_bind_host_thread_to_device
    reports the GPU's NUMA node or -1 where the platform hides the topology -- both are valid, an error status is not."""
    import ctypes as C
    import os
    from qrkit_b200 import capi
    L = capi.lib()
    nb, r, c = 4096, 8, 4
    vals = uniform_blocks(nb, r, c); b = vector(nb * r, seed=77)
    s = qk.BlockDiagonalSparseQR(pivoting=0)
    x_pageable = s.compute_solve(qk.SparseBlockDiagonal(vals, block_rows=r, block_cols=c), b).copy()

    saved = os.sched_getaffinity(0)
    try:
        node, bound = C.c_int32(-7), C.c_int32(-7)
        capi.check(L.qrk_bind_host_thread_to_device(0, C.byref(node), C.byref(bound)))
        assert node.value >= -1 and bound.value >= 0
        ptrs = []
        for nbytes in (vals.nbytes, b.nbytes, nb * c * 8):
            p = C.c_void_p()
            capi.check(L.qrk_host_alloc(C.byref(p), nbytes, 0))
            assert p.value
            ptrs.append(p)
        rt = C.CDLL("libcudart.so.12")                        # page-locked: cudaHostGetFlags only succeeds on registered memory
        rt.cudaHostGetFlags.argtypes = [C.POINTER(C.c_uint), C.c_void_p]
        for p in ptrs:
            flags = C.c_uint(0)
            assert rt.cudaHostGetFlags(C.byref(flags), p) == 0
        assert rt.cudaHostGetFlags(C.byref(flags), vals.ctypes.data_as(C.c_void_p)) != 0      # a pageable array is not
        rt.cudaGetLastError()                                   # (clear the expected error: torch shares this runtime instance)
        hv = np.ctypeslib.as_array(C.cast(ptrs[0], C.POINTER(C.c_double)), shape=(vals.size,))
        hb = np.ctypeslib.as_array(C.cast(ptrs[1], C.POINTER(C.c_double)), shape=(b.size,))
        hx = np.ctypeslib.as_array(C.cast(ptrs[2], C.POINTER(C.c_double)), shape=(nb * c,))
        hv[:] = vals; hb[:] = b; hx[:] = 0
        capi.check(L.qrk_compute_solve(s._h, ptrs[0], ptrs[1], ptrs[2], capi.QRK_HOST), s._h)
        assert np.array_equal(hx, x_pageable)
        del hv, hb, hx
        for p in ptrs:
            capi.check(L.qrk_host_free(p))
        assert L.qrk_host_alloc(None, 8, 0) == capi.QRK_STATUS_INVALID_ARGUMENT
    finally:
        os.sched_setaffinity(0, saved)
