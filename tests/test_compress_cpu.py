"""CPU: the compressed-catalogue weights (piecewise Chebyshev interpolation of the per-source term in log10 flux)."""
import numpy as np

from lumfuncmcmc_b200.compress import chebyshev_nodes, compress_sources, lagrange_basis


def _t(g, alpha, lgF, fcmin=0.1):
    """ln(modified Fleming completeness) as a function of g = log10 flux (reference VmaxLumFunc.py:118-126, 141, 164-167)."""
    n = alpha * (g - lgF)
    fc = 0.5 * (1.0 + n / np.sqrt(1.0 + n * n))
    a = (2.0 * fcmin - 1.0) ** 2
    ftau = 10 ** lgF * 10 ** (-np.sqrt(abs(a / (1.0 - a)) * alpha ** -2.0))
    return np.log(fc) / (1.0 - np.exp(-10 ** g / ftau))


def test_lagrange_basis_is_a_partition_of_unity_and_exact_at_nodes():
    nodes, bary = chebyshev_nodes(12)
    u = np.concatenate([np.linspace(-1, 1, 1001), nodes])
    L = lagrange_basis(u, nodes, bary)
    assert np.allclose(L.sum(axis=1), 1.0, atol=1e-14)
    assert np.array_equal(L[-12:], np.eye(12))
    # reproduces polynomials of degree < 12 exactly
    p = np.polynomial.Polynomial(np.random.default_rng(0).normal(size=12))
    assert np.allclose(L @ p(nodes), p(u), atol=1e-12)


def test_weighted_pseudo_sources_reproduce_the_source_sum_for_every_prior_alpha():
    rng = np.random.default_rng(1)
    g = -16.9 + rng.pareto(1.5, 400000) * 0.15
    g = g[g < -13.5]
    n = len(g)
    fi = np.array([0, n // 3, n // 3, n])                     # the middle field is empty
    xi, w, cfi = compress_sources(g, fi, alpha_max=7.0)
    assert cfi[0] == 0 and cfi[-1] == len(xi) and cfi[1] == cfi[2] and len(xi) < 6000
    assert abs(w[:cfi[1]].sum() - (n // 3)) < 1e-6                # weights of a field add up to its source count
    for alpha, lgF in ((4.56, -16.55), (7.0, -16.3), (1.0, -16.9), (7.0, -16.95), (2.5, -16.0)):
        for k in (0, 2):
            exact = _t(g[fi[k]:fi[k + 1]], alpha, lgF).sum()
            approx = (w[cfi[k]:cfi[k + 1]] * _t(xi[cfi[k]:cfi[k + 1]], alpha, lgF)).sum()
            assert abs(exact - approx) <= 1e-12 * abs(exact), (alpha, lgF, k, exact, approx)


def test_single_source_and_degenerate_fields():
    xi, w, cfi = compress_sources(np.array([-16.2]), np.array([0, 1]), alpha_max=7.0)
    assert len(xi) == 12 and abs(w.sum() - 1.0) < 1e-14
    exact = _t(np.array([-16.2]), 4.0, -16.5).sum()
    assert abs((w * _t(xi, 4.0, -16.5)).sum() - exact) < 1e-12 * abs(exact)
    xi, w, cfi = compress_sources(np.zeros(0), np.array([0, 0]), alpha_max=7.0)
    assert len(xi) == 0 and list(cfi) == [0, 0]


def test_z_model_pseudo_sources_reproduce_the_exponential_sum():
    """sum_i 10**(lum_i - L*(z_i)) with a quadratic L*(z): weighted redshift nodes against the direct sum, for gentle and
    for the steepest admitted evolution."""
    from lumfuncmcmc_b200.compress import compress_sources_z
    rng = np.random.default_rng(2)
    n = 300000
    z = rng.uniform(1.16, 1.90, n)
    lum = 41.0 + rng.pareto(1.5, n) * 0.3
    lum = np.minimum(lum, 44.5)
    fi = np.array([0, n // 4, n])
    slope_max = 60.0
    xi, v, cfi = compress_sources_z(z, lum, fi, slope_max)
    assert len(xi) < 5000 and cfi[-1] == len(xi)
    for (aL, bL, cL) in ((0.0, 0.3, 42.0), (-0.9, 3.4, 39.5), (20.0, -60.0, 87.0), (-19.0, 58.0, 0.0)):
        for k in (0, 1):
            sl = slice(fi[k], fi[k + 1])
            zz = z[sl]
            assert np.max(np.abs(2 * aL * zz + bL)) <= slope_max
            exact = np.sum(10 ** (lum[sl] - (aL * zz * zz + bL * zz + cL)))
            xs = xi[cfi[k]:cfi[k + 1]]
            approx = np.sum(v[cfi[k]:cfi[k + 1]] * 10 ** (-(aL * xs * xs + bL * xs + cL)))
            assert abs(exact - approx) <= 1e-12 * abs(exact), (aL, bL, cL, k, exact, approx)
