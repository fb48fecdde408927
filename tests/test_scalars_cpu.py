"""CPU model of householder_scalars (qrkit_b200/csrc/common.cuh): the same operation sequence in float64 with the MUFU
seeds modelled as 20-bit approximations (the measured accuracy of rsqrt.approx / rcp.approx on B200, tools/seed_accuracy.cu)
and fma modelled in extended precision.  Checks the claim of DESIGN.md §2: beta, 1/(x0 - beta) and tau agree with the IEEE
sequence Eigen's makeHouseholder runs (sqrt, two divisions) to a few ulp, for both signs of x0 and 12 decades of scale."""
import numpy as np


def _fma(a, b, c):
    return (np.longdouble(a) * np.longdouble(b) + np.longdouble(c)).astype(np.float64)


def _seed(x, bits=20):
    """a `bits`-bit approximation of x: what the MUFU unit returns"""
    m, e = np.frexp(x)
    return np.ldexp(np.round(m * 2.0 ** bits) / 2.0 ** bits, e)


def householder_scalars_model(x0, tail_sq):
    s = _fma(x0, x0, tail_sq)
    y0 = _seed(1.0 / np.sqrt(s))
    ax = np.abs(x0)
    d0 = _fma(s, y0, ax)
    r0 = _seed(1.0 / d0)
    h = 0.5 * s
    y = y0
    for _ in range(2):
        t = y * y
        e = _fma(-h, t, 0.5)
        y = _fma(y, e, y)
    nrm = s * y
    nrm = _fma(_fma(-nrm, nrm, s), 0.5 * y, nrm)
    d = ax + nrm
    e = _fma(-d, r0, 1.0)
    r = _fma(r0, e, r0)
    e = _fma(-d, r, 1.0)
    r = _fma(r, e, r)
    neg = x0 < 0.0
    beta = np.where(neg, nrm, -nrm)
    inv = np.where(neg, -r, r)
    tau = d * y
    return beta, inv, tau


def test_householder_scalars_model_matches_the_ieee_sequence():
    rng = np.random.default_rng(7)
    n = 200_000
    scale = 10.0 ** rng.uniform(-6, 6, n)
    x0 = rng.uniform(-1, 1, n) * scale
    tail_sq = (rng.uniform(0.0, 1.0, n) * scale) ** 2 * rng.integers(1, 17, n)
    tail_sq = np.maximum(tail_sq, 1e-300)
    beta, inv, tau = householder_scalars_model(x0, tail_sq)
    # Eigen makeHouseholder: beta = -sign(x0) sqrt(x0^2 + tailSq); essential = tail / (x0 - beta); tau = (beta - x0) / beta
    L = np.longdouble
    nrm = np.sqrt(L(x0) * L(x0) + L(tail_sq))
    beta_ref = np.where(x0 >= 0, -nrm, nrm)
    inv_ref = 1.0 / (L(x0) - beta_ref)
    tau_ref = (beta_ref - L(x0)) / beta_ref
    ulp = np.finfo(np.float64).eps
    assert np.max(np.abs((L(beta) - beta_ref) / beta_ref)) <= 2 * ulp
    assert np.max(np.abs((L(inv) - inv_ref) / inv_ref)) <= 3 * ulp
    assert np.max(np.abs((L(tau) - tau_ref) / tau_ref)) <= 3 * ulp
    assert np.all((tau >= 1.0 - 4 * ulp) & (tau <= 2.0 + 4 * ulp))          # tau = 1 + |x0| / norm
