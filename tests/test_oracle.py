"""CPU tests of the oracle (oracle/): the restatement is checked against
  (1) LAPACK (scipy dgeqrf / dgeqp3): R up to row sign, pivot order, and numpy lstsq for x;
  (2) the reference's own test properties (test/test-qrkit.cpp:201-203, 251-255, 289) at the
      north-star tolerances (||QR-A||/||A|| <= 1e-13, x within 1e-10) instead of the reference's 1e-6;
  (3) the exact-integer index rules of BlockDiagonalSparseQR.h:455-479, 519-521.
The reference holds no golden vectors, so value parity is pinned through these only."""
import numpy as np
import pytest
import scipy.linalg as sla

from helpers import (SEED_A, blocks_to_dense, blocks_to_sparse, dense_border, ellipse_problem,
                     overlapping_banded_matrix, rel, sign_normalize_rows, synth, uniform_blocks, vector)

SHAPES = [(2, 1), (7, 2), (8, 4), (4, 4), (16, 8), (32, 16), (128, 64), (60, 50)]


def _rand(r, c, seed=1, lo=0.5, hi=5.0):
    return uniform_blocks(1, r, c, seed=seed, lo=lo, hi=hi).reshape(c, r).T.copy()


@pytest.mark.parametrize("r,c", SHAPES)
@pytest.mark.parametrize("lo,hi", [(0.5, 5.0), (-1.0, 1.0)])
def test_householder_qr_vs_lapack(oracle, r, c, lo, hi):
    A = _rand(r, c, seed=11, lo=lo, hi=hi)
    packed, tau = oracle.householder_qr(A)
    R = np.triu(packed)[:c, :]
    Rl = sla.qr(A, mode="r")[0][:c, :]
    assert rel(sign_normalize_rows(R), sign_normalize_rows(Rl)) <= 1e-13
    Q = oracle.householder_q(packed, tau)
    assert rel(Q @ np.triu(packed), A) <= 1e-13
    assert rel(Q.T @ Q, np.eye(r)) <= 1e-13
    # Eigen's sign convention: beta = -sign(x0) * norm, so for positive inputs the first diagonal entry is < 0
    if lo > 0:
        assert packed[0, 0] < 0


@pytest.mark.parametrize("r,c", SHAPES)
@pytest.mark.parametrize("lo,hi", [(0.5, 5.0), (-1.0, 1.0)])
def test_colpiv_qr_vs_lapack(oracle, r, c, lo, hi):
    A = _rand(r, c, seed=12, lo=lo, hi=hi)
    packed, tau, perm, nz = oracle.colpiv_qr(A)
    Rl, Pl = sla.qr(A, mode="r", pivoting=True)
    assert np.array_equal(perm, Pl.astype(np.int32))          # same pivot order as dgeqp3
    R = np.triu(packed)[:c, :]
    assert rel(sign_normalize_rows(R), sign_normalize_rows(Rl[:c, :])) <= 1e-13
    Q = oracle.householder_q(packed, tau)
    assert rel(Q @ np.triu(packed), A[:, perm]) <= 1e-13
    assert nz == min(r, c)


def test_colpiv_many_small_blocks_vs_lapack(oracle):
    """Pivot-order parity over many random blocks (SURVEY §7 hard part 2)."""
    for (r, c) in [(7, 2), (8, 4)]:
        vals = uniform_blocks(2000, r, c, seed=99)
        for i in range(2000):
            A = vals[i * r * c:(i + 1) * r * c].reshape(c, r).T
            _, _, perm, _ = oracle.colpiv_qr(A)
            _, Pl = sla.qr(A, mode="r", pivoting=True)
            assert np.array_equal(perm, Pl.astype(np.int32))


def test_degenerate_columns(oracle):
    # zero tail -> tau = 0, beta = x0, no sign flip (makeHouseholder, tailSqNorm <= DBL_MIN branch)
    A = np.array([[3.0, 1.0], [0.0, 2.0], [0.0, 0.0]])
    packed, tau = oracle.householder_qr(A)
    assert tau[0] == 0.0 and packed[0, 0] == 3.0
    # 1x1 block
    packed, tau = oracle.householder_qr(np.array([[-2.5]]))
    assert tau[0] == 0.0 and packed[0, 0] == -2.5
    # exact tie: first maximum wins
    A = np.array([[1.0, 1.0, 2.0], [2.0, 2.0, 1.0], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]])
    _, _, perm, _ = oracle.colpiv_qr(A)
    assert perm[0] == 0


@pytest.mark.parametrize("r,c", [(7, 4), (12, 4), (72, 40)])
def test_block_t_factor(oracle, r, c):
    A = _rand(r, c, seed=13)
    packed, tau = oracle.householder_qr(A)
    V = np.tril(packed, -1) + np.eye(r, c)
    T = oracle.block_t_factor(V, tau)
    Q = oracle.householder_q(packed, tau)
    assert np.allclose(np.tril(T, -1), 0.0)
    assert rel(np.eye(r) - V @ T @ V.T, Q) <= 1e-13          # H_0..H_{n-1} = I - V T V^T
    # QRKit stores -T: Q = I + Y(-T)Y^T and Q^T = I + Y(-T)^T Y^T (BandedBlockedSparseQR.h:477)
    assert rel(np.eye(r) + V @ (-T).T @ V.T, Q.T) <= 1e-13


def _bd_case(nb, r, c, seed=SEED_A):
    vals = uniform_blocks(nb, r, c, seed=seed)
    br = np.full(nb, r, dtype=np.int32)
    bc = np.full(nb, c, dtype=np.int32)
    return vals, br, bc


@pytest.mark.parametrize("colpiv", [True, False])
def test_block_diagonal_reference_test0(oracle, colpiv):
    """test-qrkit.cpp:167-206 — 256 blocks 7x2 (1792 x 512)."""
    nb, r, c = 256, 7, 2
    vals, br, bc = _bd_case(nb, r, c)
    A = blocks_to_dense(vals, br, bc)
    qr = oracle.BlockDiagonalOracle(br, bc, vals, colpiv=colpiv)
    assert qr.info == 0 and qr.rank == nb * c
    Q = qr.matrixQ().toarray()
    R = qr.matrixR().toarray()
    P = qr.colsPermutation()
    AP = A[:, P]
    assert rel(Q @ R, AP) <= 1e-13
    assert rel(Q.T @ AP, R) <= 1e-13
    assert rel(Q.T @ Q, np.eye(nb * r)) <= 1e-13
    x_true = vector(nb * c, SEED_A + 1)
    b = A @ x_true
    x = qr.solve(b)
    assert rel(x, x_true) <= 1e-10
    # the same through Q^T b, triangular solve, back-permutation (test-qrkit.cpp:185-195)
    y = qr.apply_qt(b)
    solved = sla.solve_triangular(R[:nb * c, :nb * c], y[:nb * c])
    back = np.zeros(nb * c)
    back[P] = solved
    assert rel(back, x_true) <= 1e-10
    assert rel(qr.apply_q(y), b) <= 1e-13


def test_block_diagonal_index_rules(oracle):
    """Bit-exact structure of Q (FullQ), R and the column permutation (BlockDiagonalSparseQR.h:455-479, 519-521)."""
    nb, r, c = 5, 7, 2
    vals, br, bc = _bd_case(nb, r, c, seed=3)
    qr = oracle.BlockDiagonalOracle(br, bc, vals, colpiv=True)
    Q = qr.matrixQ()
    R = qr.matrixR()
    C = nb * c
    assert np.array_equal(Q.outer, np.arange(nb * r + 1) * r)
    for i in range(nb):
        for j in range(r):
            row = i * r + j
            exp = list(range(i * c, i * c + c)) + list(range(C + i * (r - c), C + (i + 1) * (r - c)))
            assert list(Q.inner[Q.outer[row]:Q.outer[row + 1]]) == exp
    for i in range(nb):
        for k in range(c):
            col = i * c + k
            assert list(R.inner[R.outer[col]:R.outer[col + 1]]) == list(range(i * c, i * c + k + 1))
    P = qr.colsPermutation()
    for i in range(nb):
        assert sorted(P[i * c:(i + 1) * c]) == list(range(i * c, (i + 1) * c))
        blk = vals[i * r * c:(i + 1) * r * c].reshape(c, r).T
        _, Pl = sla.qr(blk, mode="r", pivoting=True)
        assert np.array_equal(P[i * c:(i + 1) * c] - i * c, Pl)
    assert np.array_equal(qr.rowsPermutation(), np.arange(nb * r))


def test_block_diagonal_q_format_block_diagonal(oracle):
    nb, r, c = 6, 8, 4
    vals, br, bc = _bd_case(nb, r, c, seed=5)
    A = blocks_to_dense(vals, br, bc)
    qr = oracle.BlockDiagonalOracle(br, bc, vals, colpiv=False, qformat=1)
    Q = qr.matrixQ().toarray()
    R = qr.matrixR().toarray()
    assert rel(Q @ R, A) <= 1e-13
    for i in range(nb):   # Q is block diagonal, R rows live at base_row + j (:496-500)
        assert np.count_nonzero(Q[i * r:(i + 1) * r, :i * r]) == 0
        assert np.count_nonzero(R[i * r + c:(i + 1) * r, :]) == 0


def test_block_diagonal_ragged_and_tail(oracle):
    """Variable block sizes, an uncovered zero tail (identity rows in Q, :530-533), landscape block -> InvalidInput."""
    br = np.array([3, 7, 2, 9, 4], dtype=np.int32)
    bc = np.array([1, 2, 2, 3, 4], dtype=np.int32)
    n = int((br * bc).sum())
    vals = synth(77, 0, np.arange(n), 0)
    rows = int(br.sum()) + 3
    A = blocks_to_dense(vals, br, bc, n_rows=rows)
    qr = oracle.BlockDiagonalOracle(br, bc, vals, n_rows=rows, colpiv=True)
    Q = qr.matrixQ().toarray(); R = qr.matrixR().toarray(); P = qr.colsPermutation()
    assert rel(Q @ R, A[:, P]) <= 1e-13
    assert np.array_equal(Q[-3:, -3:], np.eye(3))
    x_true = vector(int(bc.sum()), 5)
    assert rel(qr.solve(A @ x_true), x_true) <= 1e-10
    bad = oracle.BlockDiagonalOracle(np.array([2], dtype=np.int32), np.array([3], dtype=np.int32), np.ones(6))
    assert bad.info == 3  # Eigen::InvalidInput


def test_block_diagonal_lstsq(oracle):
    """True least squares (inconsistent rhs): x equals numpy.linalg.lstsq."""
    nb, r, c = 64, 8, 4
    vals, br, bc = _bd_case(nb, r, c, seed=21)
    A = blocks_to_dense(vals, br, bc)
    b = vector(nb * r, 1234)
    x = oracle.BlockDiagonalOracle(br, bc, vals, colpiv=True).solve(b)
    assert rel(x, np.linalg.lstsq(A, b, rcond=None)[0]) <= 1e-10
    out = oracle.bd_compact_uniform(nb, r, c, vals, b, colpiv=True, threads=2)
    assert rel(out["x"], x) <= 1e-12
    assert rel(oracle.bd_reference_uniform(nb, r, c, vals, b, colpiv=True)["x"], x) <= 1e-14


def test_block_banded_pattern_merge(oracle):
    """fromBlockBandedPattern + mergeBlocks (SparseQRUtils.h:274-385): 7x4 blocks, overlap 2, SuggestedBlockCols 8."""
    blocks = oracle.block_banded_pattern(1792, 512, 7, 4, 2, 8)
    assert blocks[0].tolist() == [0, 0, 21, 8]       # 3 blocks merge: rows 21 > cols 8 >= SuggestedBlockCols
    assert all(b[2] > b[3] for b in blocks)                        # portrait
    assert blocks[-1][0] + blocks[-1][2] == 1792 and blocks[-1][1] + blocks[-1][3] == 512
    # cfg 4 reading: 16x24 slabs, step 8 -> merged 48x40 blocks
    b4 = oracle.block_banded_pattern(1600, 816, 16, 24, 16, 2)
    assert b4[0].tolist() == [0, 0, 48, 40]


@pytest.mark.parametrize("overlap_entry", [False, True])
def test_banded_reference_tests(oracle, overlap_entry):
    """test-qrkit.cpp:208-258 (tests 1/2): Q R = A, Q^T A = R with Q formed explicitly, x recovered."""
    n_params, n_res = 128, 448
    A = overlapping_banded_matrix(n_params, n_res, overlap_entry=overlap_entry)
    blocks = oracle.block_banded_pattern(n_res, n_params, 7, 4 if overlap_entry else 2, 2 if overlap_entry else 0, 8)
    qr = oracle.BandedOracle(A, blocks)
    Ad = A.toarray()
    R = qr.matrixR().toarray()
    I = np.eye(n_res)
    Q = np.column_stack([qr.apply_q(I[:, j], False) for j in range(n_res)])
    Qt = np.column_stack([qr.apply_q(I[:, j], True) for j in range(n_res)])
    assert rel(Q @ R, Ad) <= 1e-13
    assert rel(Q.T @ Ad, R) <= 1e-13
    assert rel(Qt.T @ R, Ad) <= 1e-13
    assert rel(Qt @ Ad, R) <= 1e-13
    assert np.allclose(np.tril(R, -1), 0.0)
    x_true = vector(n_params, 3)
    assert rel(qr.solve(Ad @ x_true), x_true) <= 1e-10
    Rl = sla.qr(Ad, mode="r")[0][:n_params]
    assert rel(sign_normalize_rows(R[:n_params]), sign_normalize_rows(Rl)) <= 1e-12


@pytest.mark.parametrize("right_kind", [0, 1])
def test_block_angular_banded_left(oracle, right_kind):
    """test-qrkit.cpp:260-327 (tests 4/5) at reduced size: banded left + dense border, x recovered."""
    n_params, n_res, m2 = 128, 448, 24
    A1 = overlapping_banded_matrix(n_params, n_res)
    J2 = dense_border(n_res, m2)
    blocks = oracle.block_banded_pattern(n_res, n_params, 7, 4, 2, 8)
    qr = oracle.BlockAngularOracle(J2, A_csc=A1, blocks=blocks, right_kind=right_kind, panel=2)
    A = np.hstack([A1.toarray(), J2])
    x_true = vector(n_params + m2, 9)
    b = A @ x_true
    assert rel(qr.solve(b), x_true) <= 1e-10
    R = qr.matrixR().toarray()
    P = qr.colsPermutation()
    Rl = sla.qr(A[:, P], mode="r")[0][:n_params + m2]
    assert rel(sign_normalize_rows(R[:n_params + m2]), sign_normalize_rows(Rl)) <= 1e-12
    y = qr.apply_qt(b)
    assert rel(np.linalg.norm(y), np.linalg.norm(b)) <= 1e-13 or abs(np.linalg.norm(y) - np.linalg.norm(b)) <= 1e-10


def test_block_angular_ellipse(oracle):
    """cfg 1/3 shape: N blocks 2x1 + dense 2N x 5 border, ColPiv left and right (bench_sparse_qr_extra.cpp:153-174)."""
    n = 500
    J1, J2, rhs = ellipse_problem(n)
    br = np.full(n, 2, dtype=np.int32); bc = np.ones(n, dtype=np.int32)
    qr = oracle.BlockAngularOracle(J2, br=br, bc=bc, values=J1, left_colpiv=True, right_kind=0)
    A = np.hstack([blocks_to_dense(J1, br, bc), J2])
    b = vector(2 * n, 17)
    x = qr.solve(b)
    assert rel(x, np.linalg.lstsq(A, b, rcond=None)[0]) <= 1e-9
    R = qr.matrixR(); P = qr.colsPermutation()
    Rd = R.toarray()
    assert rel(Rd.T @ Rd, A[:, P].T @ A[:, P]) <= 1e-12           # R^T R = (AP)^T (AP)
    # structure (makeR, BlockAngularSparseQR.h:285-308): n diagonal entries, then m2 columns of n + (c+1) entries
    assert R.outer[n] == n and np.array_equal(R.inner[:n], np.arange(n))
    for c in range(5):
        assert R.outer[n + c + 1] - R.outer[n + c] == n + c + 1
    assert np.array_equal(P[:n], np.arange(n)) and sorted(P[n:]) == list(range(n, n + 5))
    out = oracle.angular_reference_uniform(n, 2, 1, J1, J2, b)
    assert rel(out["x"], x) <= 1e-14


@pytest.mark.parametrize("br,bc,ov", [(16, 24, 16), (7, 4, 2), (7, 2, 0), (8, 8, 4), (12, 8, 4), (4, 6, 4)])
@pytest.mark.parametrize("nb", [1, 2, 3, 5, 11, 40])
def test_oracle_banded_q_geometries(oracle, br, bc, ov, nb):
    """Where the reference's banded Q is an exact factor and where it is not (a reference limitation the oracle restates
    faithfully, BandedBlockedSparseQR.h:497-500): x = R^-1 (Q^T b)[0:n] equals LAPACK's least-squares solution exactly for the
    geometries helpers.oracle_banded_q_is_exact names, and is visibly wrong (> 1e-3) for the others; R equals LAPACK's up to
    row signs for ALL of them."""
    from helpers import oracle_banded_q_is_exact, reference_style_windows, sign_normalize_rows, slabs_to_sparse
    slabs = uniform_blocks(nb, br, bc)
    A = slabs_to_sparse(slabs, nb, br, bc, ov)
    if A.shape[0] < A.shape[1]:
        pytest.skip("fewer rows than columns")
    windows = reference_style_windows(nb, br, bc, ov, 2)
    ref = oracle.BandedOracle(A, windows)
    Ad = A.toarray()
    R = ref.matrixR().toarray()[:A.shape[1], :]
    assert rel(sign_normalize_rows(R), sign_normalize_rows(np.linalg.qr(Ad, mode="r"))) <= 1e-12
    b = vector(A.shape[0], seed=3)
    err = rel(ref.solve(b), np.linalg.lstsq(Ad, b, rcond=None)[0])
    if oracle_banded_q_is_exact(nb, windows):
        assert err <= 1e-10
    else:
        assert err > 1e-3
