"""The design choice behind the block-angular border (difference 3): column-pivoted QR of the tall residual
(what the reference runs, BlockAngularSparseQR.h:368) and column-pivoted QR of the m2 x m2 triangle of an UNPIVOTED QR (TSQR
or blocked compact-WY first stage — what the device runs) give the same permutation and the same R up to row signs: the
triangle has the column norms and inner products of the tall matrix.  Checked with LAPACK's dgeqp3 (same pivot rule as Eigen's
ColPivHouseholderQR on generic data) on the shapes of the reference's tests."""
import numpy as np
import pytest
import scipy.linalg as sla


@pytest.mark.parametrize("n,m2,seed", [(2000, 5, 1), (5120, 384, 2), (600, 24, 3), (480, 96, 4)])
def test_colpiv_of_the_triangle_equals_colpiv_of_the_tall_matrix(n, m2, seed):
    rng = np.random.default_rng(seed)
    A = rng.uniform(0.5, 5.0, (n, m2))
    R_tall, P_tall = sla.qr(A, mode="r", pivoting=True)
    Rt = np.linalg.qr(A, mode="r")                         # unpivoted first stage
    R_tri, P_tri = sla.qr(Rt, mode="r", pivoting=True)
    assert np.array_equal(P_tall, P_tri)
    a, b = np.abs(R_tall[:m2, :]), np.abs(R_tri[:m2, :])
    assert np.linalg.norm(a - b) <= 1e-12 * np.linalg.norm(a)
