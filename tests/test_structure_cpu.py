"""Host-side pattern analysis against the reference's own known answers (test/test-utils.cpp:182-274: exact block lists
after AsBandedAsPossible row ordering and BlockBandedMatrixInfo detection) and against the oracle's restatement of the
pattern functions.  Integer work: every comparison is bit-exact.  No GPU needed (the functions are host code in the
C-ABI library, as SparseQROrdering / SparseQRUtils are host code in the reference)."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import overlapping_banded_matrix, synth


@pytest.fixture(scope="module")
def st():
    from qrkit_b200 import structure
    return structure


def block_diagonal_7x2(num_params, num_residuals):
    """generate_block_diagonal_matrix (test/test-utils.cpp:60-98 / test-qrkit.cpp:101-131): block row i (7 rows) holds
    columns 2i, 2i+1 dense."""
    rows, cols = [], []
    for i in range(num_params):
        for j in range(2 * i, min(2 * i + 2, num_params)):
            for k in range(7):
                rows.append(7 * i + k); cols.append(j)
    rows = np.array(rows); cols = np.array(cols)
    keep = rows < num_residuals
    rows, cols = rows[keep], cols[keep]
    return sp.csc_matrix((synth(1, 0, rows, cols), (rows, cols)), shape=(num_residuals, num_params))


def shuffle_rows(A, seed):
    rng = np.random.default_rng(seed)
    p = rng.permutation(A.shape[0])
    return A.tocsr()[p, :]


NUM_VARS = 256
NUM_PARAMS = 2 * NUM_VARS
NUM_RES = 7 * NUM_VARS          # numVars*3 + numVars + numVars*3 (test-utils.cpp:361-363)


def _order_then_detect(st, A):
    perm, has = st.as_banded_as_possible(A)
    Ar = A.tocsr()
    if has:
        inv = np.empty_like(perm); inv[perm] = np.arange(len(perm), dtype=np.int32)
        Ar = Ar[inv, :]
    return st.detect_blocks(Ar)[0], perm, has


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_blockdiag_permuted(st, seed):
    """test_blockdiag_permuted (test-utils.cpp:182-209): 256 blocks at (7i, 2i) of 7x2 after undoing a row shuffle."""
    A = shuffle_rows(block_diagonal_7x2(NUM_PARAMS, NUM_RES), seed)
    blocks, perm, has = _order_then_detect(st, A)
    assert has
    assert np.array_equal(np.sort(perm), np.arange(NUM_RES))
    i = np.arange(256)
    assert np.array_equal(blocks, np.stack([7 * i, 2 * i, np.full(256, 7), np.full(256, 2)], axis=1))


@pytest.mark.parametrize("seed", [0, 3])
def test_overlapping_permuted(st, seed):
    """test_overlapping_permuted (test-utils.cpp:211-252): 255 blocks, 7x4 at (7i, 2i), the last one 14x4."""
    A = shuffle_rows(overlapping_banded_matrix(NUM_PARAMS, NUM_RES), seed)
    blocks, _, has = _order_then_detect(st, A)
    assert has and len(blocks) == 255
    i = np.arange(255)
    expect = np.stack([7 * i, 2 * i, np.full(255, 7), np.full(255, 4)], axis=1)
    expect[-1, 2] = 14
    assert np.array_equal(blocks, expect)


def test_blockdiag_vertperm_diag(st):
    """test_blockdiag_vertperm_diag (test-utils.cpp:145-180, 254-274): sqrt(lambda) I interleaved below the last entry of each
    column gives 256 blocks of 9x2 at (9i, 2i), found WITHOUT row reordering."""
    A = block_diagonal_7x2(NUM_PARAMS, NUM_RES).tocsc()
    n_res, n_par = A.shape
    new_row = np.empty(n_res + n_par, dtype=np.int64)
    curr = 0
    for c in range(n_par):
        last = A.indices[A.indptr[c]:A.indptr[c + 1]].max()
        while curr <= last + c:
            new_row[curr - c] = curr
            curr += 1
        new_row[n_res + c] = curr
        curr += 1
    stacked = sp.vstack([A, np.sqrt(1e-3) * sp.identity(n_par)]).tocoo()
    out = sp.csr_matrix((stacked.data, (new_row[stacked.row], stacked.col)), shape=(n_res + n_par, n_par))
    blocks, _ = st.detect_blocks(out)
    i = np.arange(256)
    assert np.array_equal(blocks, np.stack([9 * i, 2 * i, np.full(256, 9), np.full(256, 2)], axis=1))
    perm, has = st.as_banded_as_possible(out)
    assert not has and np.array_equal(perm, np.arange(n_res + n_par))


def test_already_ordered_matrix_keeps_its_rows(st):
    A = block_diagonal_7x2(64, 224)
    perm, has = st.as_banded_as_possible(A)
    assert not has and np.array_equal(perm, np.arange(224))


def test_stable_sort_keeps_row_order_inside_a_band(st):
    """std::stable_sort by band start only (SparseQROrdering.h:87-118): rows sharing a start keep their relative order."""
    A = sp.csr_matrix(np.array([[0, 0, 1, 1], [1, 1, 0, 0], [0, 0, 1, 0], [1, 0, 0, 0], [0, 0, 0, 0]], dtype=float))
    perm, has = st.as_banded_as_possible(A)
    assert has
    assert perm.tolist() == [2, 0, 3, 1, 4]          # starts 2,0,2,0,4 -> new order rows [1,3,0,2,4]


def test_column_density(st):
    A = sp.csc_matrix(np.array([[1, 1, 0, 1], [1, 0, 0, 1], [1, 0, 0, 0]], dtype=float))
    assert st.column_density(A).tolist() == [3, 1, 0, 2]   # nnz 3,1,0,2 -> ascending, stable: cols [2,1,3,0]


@pytest.mark.parametrize("rows,cols,br,bc,ov,sug", [(1792, 512, 7, 4, 2, 2), (1600, 816, 16, 24, 16, 2), (1600, 816, 16, 24, 16, 8),
                                                     (96, 28, 8, 8, 4, 2), (70, 20, 7, 2, 0, 2), (64, 30, 4, 6, 4, 2)])
def test_pattern_functions_match_the_oracle(st, oracle, rows, cols, br, bc, ov, sug):
    ours = st.block_banded_pattern(rows, cols, br, bc, ov, sug)
    ref = oracle.block_banded_pattern(rows, cols, br, bc, ov, sug)
    assert np.array_equal(ours, np.asarray(ref, dtype=np.int32).reshape(-1, 4))


def test_block_diagonal_pattern(st):
    b = st.block_diagonal_pattern(1792, 512, 7, 2)
    i = np.arange(256)
    assert np.array_equal(b, np.stack([7 * i, 2 * i, np.full(256, 7), np.full(256, 2)], axis=1))


def test_extract_blocks_and_from_sparse_matrix(st):
    A0 = block_diagonal_7x2(40, 140)
    A = shuffle_rows(A0, 5)
    vals, br, bc, blocks, perm, has = st.from_sparse_matrix(A)
    assert has and np.all(br == 7) and np.all(bc == 2) and len(br) == 20
    inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
    PA = A.tocsr()[inv, :].toarray()
    o = 0
    for (r0, c0, nr, nc) in blocks:
        assert np.array_equal(vals[o:o + nr * nc].reshape(nc, nr).T, PA[r0:r0 + nr, c0:c0 + nc])
        o += nr * nc
    # everything outside the blocks is zero: the blocks carry the whole matrix
    assert np.count_nonzero(vals) == A.nnz


def test_empty_and_degenerate_inputs(st):
    assert len(st.detect_blocks(sp.csr_matrix((0, 5)))[0]) == 0
    assert len(st.detect_blocks(sp.csr_matrix((4, 3)))[0]) == 0          # all rows empty: no block
    b, _ = st.detect_blocks(sp.csr_matrix(np.ones((5, 2))))
    assert b.tolist() == [[0, 0, 5, 2]]
