"""Pins the oracle's special functions and GLL setup against independent libraries."""
import mpmath
import numpy as np
import pytest

from oracle import oracle


@pytest.mark.parametrize("a", [-1.0, -0.101, -0.5 + 1.5 * 0.266, 0.0, 0.899, 2.899])
def test_gamma_upper_matches_mpmath(a):
    oracle.build()
    for x in [0.05, 0.068, 0.3, 1.0, 1.817, 2.5, 7.0]:
        ref = float(mpmath.gammainc(a, x))
        got = oracle.gamma_upper(a, x)
        assert got == pytest.approx(ref, rel=1e-12), (a, x)


def test_expint_E1_matches_mpmath():
    for x in [1e-3, 0.068, 0.5, 1.0, 1.5, 3.0, 10.0]:
        assert oracle.expint_E1(x) == pytest.approx(float(mpmath.e1(x)), rel=1e-13)


@pytest.mark.parametrize("Nq", [2, 3, 5, 8])
def test_gll_nodes_weights_derivative(Nq):
    x, w, D = oracle.gll(Nq)
    N = Nq - 1
    # nodes are ±1 and the roots of P_N'
    dPN = np.polynomial.legendre.Legendre.basis(N).deriv()
    ref = np.sort(np.concatenate([[-1.0], dPN.roots().real, [1.0]])) if N > 1 else np.array([-1.0, 1.0])
    assert np.allclose(x, ref, atol=1e-14)
    assert np.isclose(w.sum(), 2.0)
    # quadrature is exact up to degree 2N-1
    for deg in range(2 * N):
        exact = 0.0 if deg % 2 else 2.0 / (deg + 1)
        assert np.isclose(np.dot(w, x**deg), exact, atol=1e-13)
    # D differentiates polynomials of degree <= N exactly; rows sum to 0
    assert np.allclose(D.sum(axis=1), 0, atol=1e-12)
    for deg in range(1, N + 1):
        assert np.allclose(D @ x**deg, deg * x ** (deg - 1), atol=1e-11)
