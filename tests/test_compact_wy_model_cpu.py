"""CPU model (numpy) of the compact-WY algebra of qrkit_b200/csrc/dense_blocked.cuh (and bd_wy.cuh): the panel's reflectors in
Eigen's makeHouseholder convention, the triangular factor from G = V^T V by the column recurrence the panel kernel runs, and the
trailing update A2 <- A2 - V (T^T (V^T A2)) the DMMA kernels apply — checked against the product of the individual reflectors
(the form BlockedThinQRBase::updateMat applies one column at a time, src/QRKit/BlockedThinQRBase.h:308-333)."""
import numpy as np
import pytest


def panel(A):
    """unpivoted Householder QR of the panel A (rows x pw): returns V (unit lower trapezoidal), tau, R"""
    A = A.copy()
    rows, pw = A.shape
    V = np.zeros((rows, pw)); tau = np.zeros(pw)
    for c in range(pw):
        x0, tail = A[c, c], A[c + 1:, c]
        ts = tail @ tail
        if ts <= np.finfo(float).tiny:
            t, beta, inv = 0.0, x0, 0.0
        else:
            beta = -np.copysign(np.sqrt(x0 * x0 + ts), x0) if x0 != 0 else -np.sqrt(ts)
            inv = 1.0 / (x0 - beta); t = (beta - x0) / beta
        v = np.concatenate([[1.0], tail * inv])
        w = t * (v @ A[c:, c + 1:])
        A[c:, c + 1:] -= np.outer(v, w)
        A[c, c] = beta; A[c + 1:, c] = 0.0
        V[c:, c] = v; tau[c] = t
    return V, tau, A[:pw, :]


def t_factor(V, tau):
    """T[0:j, j] = -tau_j T[0:j, 0:j] G[0:j, j], T[j, j] = tau_j with G = V^T V (dense_panel_kernel, thread 0)"""
    pw = V.shape[1]
    G = V.T @ V
    T = np.zeros((pw, pw))
    for j in range(pw):
        T[j, j] = tau[j]
        T[:j, j] = -tau[j] * (T[:j, :j] @ G[:j, j])
    return T


@pytest.mark.parametrize("rows,pw,ntrail", [(40, 8, 13), (9, 8, 5), (300, 8, 24), (50, 3, 7)])
def test_compact_wy_equals_the_product_of_reflectors(rows, pw, ntrail):
    rng = np.random.default_rng(rows + pw)
    A = rng.uniform(0.5, 5.0, (rows, pw + ntrail))
    V, tau, R = panel(A[:, :pw])
    T = t_factor(V, tau)
    Q = np.eye(rows)
    for c in range(pw):
        Q = Q @ (np.eye(rows) - tau[c] * np.outer(V[:, c], V[:, c]))                 # H_0 H_1 ... H_{pw-1}
    assert np.allclose(np.eye(rows) - V @ T @ V.T, Q, atol=1e-13)
    A2 = A[:, pw:]
    W = V.T @ A2                                                                      # dense_wy_w_kernel
    X = -(T.T @ W)                                                                    # dense_wy_apply_kernel
    assert np.allclose(A2 + V @ X, Q.T @ A2, atol=1e-12)                              # = the reflectors applied one by one
    assert np.allclose(np.abs(R), np.abs(np.linalg.qr(A[:, :pw], mode="r")), atol=1e-12)
