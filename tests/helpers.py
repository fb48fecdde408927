"""Input generators shared by the CPU and GPU tests.

Shapes follow the reference's generators (test/test-qrkit.cpp:63-165) but values come from the shared
counter-based RNG so that CPU and GPU see identical inputs on any platform."""
import numpy as np
import scipy.sparse as sp

SEED_A = 0x51524B49  # "QRKI"
_MASK = (1 << 64) - 1


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def synth(seed, block, row, col, lo=0.5, hi=5.0):
    """U[lo, hi) from splitmix64(seed ^ (block<<20) ^ (row<<10) ^ col); broadcasts over numpy arrays."""
    block = np.asarray(block, dtype=np.uint64)
    row = np.asarray(row, dtype=np.uint64)
    col = np.asarray(col, dtype=np.uint64)
    u = splitmix64(np.uint64(seed) ^ (block << np.uint64(20)) ^ (row << np.uint64(10)) ^ col)
    return lo + (hi - lo) * (u >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform_blocks(nb, r, c, seed=SEED_A, lo=0.5, hi=5.0, block0=0):
    """Flat block-major values, col-major inside each block: value(block i, row k, col j)."""
    i = np.arange(block0, block0 + nb, dtype=np.uint64)[:, None, None]
    j = np.arange(c, dtype=np.uint64)[None, :, None]
    k = np.arange(r, dtype=np.uint64)[None, None, :]
    return synth(seed, i, k, j, lo, hi).reshape(-1)


def vector(n, seed, lo=-1.0, hi=1.0):
    return synth(seed, np.arange(n, dtype=np.uint64), 0, 0, lo, hi)


def blocks_to_dense(values, br, bc, n_rows=None, n_cols=None):
    br = np.asarray(br); bc = np.asarray(bc)
    R = int(br.sum()) if n_rows is None else n_rows
    Cc = int(bc.sum()) if n_cols is None else n_cols
    A = np.zeros((R, Cc))
    o = r0 = c0 = 0
    for r, c in zip(br, bc):
        A[r0:r0 + r, c0:c0 + c] = values[o:o + r * c].reshape(c, r).T
        o += r * c; r0 += r; c0 += c
    return A


def blocks_to_sparse(values, br, bc, n_rows=None, n_cols=None):
    br = np.asarray(br); bc = np.asarray(bc)
    R = int(br.sum()) if n_rows is None else n_rows
    Cc = int(bc.sum()) if n_cols is None else n_cols
    rows, cols = [], []
    r0 = c0 = 0
    for r, c in zip(br, bc):
        jj, kk = np.meshgrid(np.arange(c), np.arange(r), indexing="ij")
        rows.append((r0 + kk).reshape(-1)); cols.append((c0 + jj).reshape(-1))
        r0 += r; c0 += c
    return sp.csc_matrix((values, (np.concatenate(rows), np.concatenate(cols))), shape=(R, Cc))


def overlapping_banded_matrix(num_params, num_residuals, seed=SEED_A, overlap_entry=True):
    """Pattern of generate_overlapping_block_diagonal_matrix (test-qrkit.cpp:63-96) without the row shuffle:
    for block row i (7 rows) columns 2i, 2i+1 dense, plus (7i+6, j+2) when j < num_params-2."""
    rows, cols = [], []
    stride = 7
    for i in range(num_params):
        for j in range(2 * i, min(2 * i + 2, num_params)):
            for k in range(7):
                rows.append(i * stride + k); cols.append(j)
            if overlap_entry and j < num_params - 2:
                rows.append(i * stride + 6); cols.append(j + 2)
    rows = np.array(rows); cols = np.array(cols)
    vals = synth(seed, 0, rows, cols)
    return sp.csc_matrix((vals, (rows, cols)), shape=(num_residuals, num_params))


def dense_border(num_residuals, m2, seed=SEED_A + 7):
    i = np.arange(num_residuals, dtype=np.uint64)[:, None]
    j = np.arange(m2, dtype=np.uint64)[None, :]
    return synth(seed, 1, i, j)


def ellipse_problem(n, noise=0.0):
    """Ellipse-fit Jacobian at the initial LM iterate (bench/bench_sparse_qr_extra.cpp:79-114, 221-282).
    Returns (J1 blocks flat (n blocks 2x1), J2 dense 2n x 5, rhs 2n)."""
    a, b, x0, y0, r = 7.5, 2.0, 17.0, 23.0, 0.23
    incr = 1.3 * np.pi / n
    t = np.arange(n) * incr
    px = x0 + a * np.cos(t) * np.cos(r) - b * np.sin(t) * np.sin(r)
    py = y0 + a * np.cos(t) * np.sin(r) + b * np.sin(t) * np.cos(r)
    pa = 0.5 * (px.max() - px.min()); pb = 0.5 * (py.max() - py.min())
    cx = 0.5 * (px.max() + px.min()); cy = 0.5 * (py.max() + py.min()); pr = 0.0
    ct, st, cr, sr = np.cos(t), np.sin(t), np.cos(pr), np.sin(pr)
    J1 = np.empty((n, 2))
    J1[:, 0] = pa * cr * st + pb * sr * ct
    J1[:, 1] = pa * sr * st - pb * cr * ct
    J2 = np.zeros((2 * n, 5))
    J2[0::2, 0] = -ct * cr; J2[0::2, 1] = st * sr; J2[0::2, 2] = -1; J2[0::2, 4] = pa * ct * sr + pb * st * cr
    J2[1::2, 0] = -ct * sr; J2[1::2, 1] = -st * cr; J2[1::2, 3] = -1; J2[1::2, 4] = -pa * ct * cr + pb * st * sr
    fx = px - (pa * ct * cr - pb * st * sr + cx)
    fy = py - (pa * ct * sr + pb * st * cr + cy)
    rhs = np.empty(2 * n); rhs[0::2] = fx; rhs[1::2] = fy
    return J1.reshape(-1), J2, rhs


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def sign_normalize_rows(R):
    """Flip rows so that the diagonal is non-negative (R is unique up to row signs)."""
    R = np.array(R, dtype=np.float64, copy=True)
    n = min(R.shape)
    for i in range(n):
        if R[i, i] < 0:
            R[i, :] = -R[i, :]
    return R


def reference_style_windows(nb, br, bc, ov, suggested=2):
    """Windows the reference would factor for nb block rows of br x bc shifted by s = bc - ov columns: consecutive block rows
    are merged until the window is portrait and at least `suggested` columns wide (mergeBlocks rule, SparseQRUtils.h:357),
    a leftover that cannot form such a window is merged into the last one (:370-383).  Rows: (idxRow, idxCol, numRows, numCols)."""
    s = bc - ov
    out, first, rows, cols = [], None, 0, 0
    for k in range(nb):
        if first is None:
            first, rows, cols = k, br, bc
        else:
            rows, cols = (k + 1 - first) * br, (k - first) * s + bc
        if rows > cols and cols >= s and cols >= suggested:
            out.append([first * br, first * s, rows, cols])
            first = None
    if first is not None:
        if out:
            last = out[-1]
            out[-1] = [last[0], last[1], last[2] + rows, first * s + cols - last[1]]
        else:
            out.append([first * br, first * s, rows, cols])
    return np.array(out, dtype=np.int32)


def slabs_to_sparse(slabs, nb, br, bc, ov):
    """The block-banded matrix of fromBlockBandedPattern (SparseQRUtils.h:274-302) from its dense slabs (column-major,
    back to back): slab k at rows [k*br, ...), columns [k*(bc-ov), ...)."""
    s = bc - ov
    S = np.asarray(slabs).reshape(nb, bc, br)
    jj, ii = np.meshgrid(np.arange(bc), np.arange(br), indexing="ij")
    rows = (np.arange(nb)[:, None, None] * br + ii[None]).reshape(-1)
    cols = (np.arange(nb)[:, None, None] * s + jj[None]).reshape(-1)
    return sp.csc_matrix((S.reshape(-1), (rows, cols)), shape=(nb * br, (nb - 1) * s + bc))


def oracle_banded_q_is_exact(nb, windows):
    """Geometries for which the reference's banded Q (and therefore the oracle's, which restates it) is an exact factor of A:
    no slabs were merged (one window per slab), or the chain has at most two windows (no merged window in the middle).  For the
    others `activeRows = bi.numRows + ...` (BandedBlockedSparseQR.h:497) undercounts the rows a merged middle window hands on in
    the 2-segment YTY layout; R is unaffected.  Enumerated in tests/test_oracle.py::test_oracle_banded_q_geometries."""
    return len(windows) == nb or len(windows) <= 2
