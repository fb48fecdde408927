"""Executable specification (numpy, CPU) of the two-phase banded factorisation of qrkit_b200/csrc/banded.cuh:
  phase 1  every group of G consecutive slabs is reduced on its own, from a zero carry, to a band triangle R_g whose last OV
           rows overlap the next group's first OV columns;
  phase 2  the OV-row triangle handed over by group g-1 is merged into R_g; the first G*S rows become final rows of R, the
           last OV rows are handed on, OV rows are annihilated.
Checked against LAPACK on the assembled matrix: R up to row signs, x, and the isometry the block-angular path relies on —
thin part and complement of Q^T b together carry the whole norm of b (the virtual zero rows add nothing)."""
import numpy as np
import pytest

from helpers import rel, sign_normalize_rows, slabs_to_sparse, uniform_blocks, vector


def two_phase(slabs, nb, br, bc, ov, G, b):
    s = bc - ov
    n_cols = (nb - 1) * s + bc
    S = np.asarray(slabs).reshape(nb, bc, br).transpose(0, 2, 1)          # nb x br x bc
    R = np.zeros((n_cols, n_cols))
    y = np.zeros(n_cols)
    comp = []
    carry = np.zeros((ov, ov + 1))                                         # [triangle | rhs] handed between groups
    for g0 in range(0, nb, G):
        n_g = min(G, nb - g0)
        wc = (n_g - 1) * s + bc                                            # columns (= pivot rows) of the group
        # ---- phase 1: the group's slabs from a ZERO carry (ov virtual zero rows on top)
        M = np.zeros((ov + n_g * br, wc + 1))
        for k in range(n_g):
            M[ov + k * br: ov + (k + 1) * br, k * s: k * s + bc] = S[g0 + k]
            M[ov + k * br: ov + (k + 1) * br, wc] = b[(g0 + k) * br:(g0 + k + 1) * br]
        Q1, _ = np.linalg.qr(M[:, :wc], mode="complete")
        T = Q1.T @ M
        Rg, comp1 = T[:wc, :], T[wc:, wc]                                  # band triangle [R_g | y_g]; annihilated rows' rhs
        comp.append(comp1)
        # ---- phase 2: merge the handed triangle into R_g
        N = np.zeros((ov + wc, wc + 1))
        N[:ov, :ov] = carry[:, :ov]; N[:ov, wc] = carry[:, ov]
        N[ov:, :] = Rg
        Q2, _ = np.linalg.qr(N[:, :wc], mode="complete")
        U = Q2.T @ N
        last = g0 + n_g >= nb
        nfinal = wc if last else n_g * s
        c0 = g0 * s
        R[c0:c0 + nfinal, c0:c0 + wc] = U[:nfinal, :wc]
        y[c0:c0 + nfinal] = U[:nfinal, wc]
        if not last:
            carry = np.hstack([U[nfinal:wc, wc - ov:wc], U[nfinal:wc, wc:wc + 1]])    # the last OV rows, their OV columns
        comp.append(U[wc:, wc])                                            # the OV annihilated rows of the merge
    return R, y, np.concatenate(comp)


@pytest.mark.parametrize("br,bc,ov", [(16, 24, 16), (7, 4, 2), (7, 2, 0), (12, 8, 4)])
@pytest.mark.parametrize("G", [1, 3, 4, 50])
def test_two_phase_model_matches_lapack(br, bc, ov, G):
    nb = 11
    slabs = uniform_blocks(nb, br, bc)
    A = slabs_to_sparse(slabs, nb, br, bc, ov).toarray()
    b = vector(nb * br, seed=3)
    R, y, comp = two_phase(slabs, nb, br, bc, ov, G, b)
    Rref = np.linalg.qr(A, mode="r")
    assert np.allclose(np.tril(R, -1), 0.0, atol=1e-12)
    assert rel(sign_normalize_rows(R), sign_normalize_rows(Rref)) <= 1e-12
    x = np.linalg.solve(R, y)
    assert rel(x, np.linalg.lstsq(A, b, rcond=None)[0]) <= 1e-11
    # isometry: thin part and complement together carry all of b
    assert abs(y @ y + comp @ comp - b @ b) <= 1e-12 * (b @ b)
    # the complement has the dimension the device path allocates: nb (ov + br - bc) + groups * ov  (+ 0: full last slab)
    groups = (nb + G - 1) // G
    assert comp.size == nb * (ov + br - bc) + groups * ov


@pytest.mark.parametrize("br,bc,ov,G", [(16, 24, 16, 3), (7, 4, 2, 4), (7, 2, 0, 2)])
def test_block_angular_on_the_extended_complement(br, bc, ov, G):
    """BlockAngularSparseQR with a banded left block as the device path runs it: Q1^T [J2 | b] by the two-phase application,
    the dense right block factored on the EXTENDED complement (it contains a few virtual zero rows; the map is an isometry, so
    the least-squares problem is unchanged), x1 = R1^-1 (y1 - Atop x2)."""
    from helpers import dense_border
    nb, m2 = 9, 5
    slabs = uniform_blocks(nb, br, bc)
    A1 = slabs_to_sparse(slabs, nb, br, bc, ov).toarray()
    n, m1 = A1.shape
    J2 = dense_border(n, m2)
    b = vector(n, seed=5)
    cols = [two_phase(slabs, nb, br, bc, ov, G, J2[:, j]) for j in range(m2)]
    R1, y1, comp_b = two_phase(slabs, nb, br, bc, ov, G, b)
    Atop = np.stack([c[1] for c in cols], axis=1)
    Abot = np.stack([c[2] for c in cols], axis=1)
    x2 = np.linalg.lstsq(Abot, comp_b, rcond=None)[0]
    x1 = np.linalg.solve(R1, y1 - Atop @ x2)
    x_ref = np.linalg.lstsq(np.hstack([A1, J2]), b, rcond=None)[0]
    assert rel(np.concatenate([x1, x2]), x_ref) <= 1e-10
