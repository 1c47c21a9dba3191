"""CPU: the ensemble sampler (host logic around the batched lnprob) on an analytic target."""
import numpy as np
import pytest

from lumfuncmcmc_b200.sampler import EnsembleSampler, integrated_time


def _gauss(mu, sig):
    def f(x):
        x = np.atleast_2d(x)
        return -0.5 * np.sum(((x - mu) / sig) ** 2, axis=1)
    return f


def test_vectorised_and_scalar_modes_are_the_same_chain():
    mu, sig = np.array([1.0, -2.0, 0.5]), np.array([0.5, 2.0, 1.0])
    f = _gauss(mu, sig)
    np.random.seed(3)
    p0 = mu + 0.1 * np.random.randn(20, 3)
    state = np.random.get_state()
    a = EnsembleSampler(20, 3, f, vectorize=True)
    a.run_mcmc(p0, 50, rstate0=state)
    b = EnsembleSampler(20, 3, lambda x: float(f(x)[0]), vectorize=False)
    b.run_mcmc(p0, 50, rstate0=state)
    assert np.array_equal(a.chain, b.chain) and np.array_equal(a.lnprobability, b.lnprobability)
    assert a.chain.shape == (20, 50, 3) and a.lnprobability.shape == (20, 50)
    assert a.ncalls == 1 + 2 * 50                      # one batched call per half-ensemble per step
    assert b.ncalls == a.ncalls


def test_recovers_gaussian_moments_and_reports_diagnostics():
    mu, sig = np.array([1.0, -2.0]), np.array([0.5, 2.0])
    np.random.seed(5)
    s = EnsembleSampler(40, 2, _gauss(mu, sig), vectorize=True)
    s.run_mcmc(mu + 0.1 * np.random.randn(40, 2), 1500, rstate0=np.random.get_state())
    flat = s.chain[:, 300:, :].reshape(-1, 2)
    assert np.all(np.abs(flat.mean(axis=0) - mu) < 0.15 * sig)
    assert np.all(np.abs(flat.std(axis=0) / sig - 1.0) < 0.1)
    acc = s.acceptance_fraction
    assert acc.shape == (40,) and 0.3 < acc.mean() < 0.9
    tau = s.acor
    assert tau.shape == (2,) and np.all(tau > 1.0) and np.all(tau < 200.0)


def test_minus_inf_proposals_are_never_accepted_and_nan_raises():
    def f(x):
        x = np.atleast_2d(x)
        lp = -0.5 * np.sum(x ** 2, axis=1)
        lp[x[:, 0] > 1.0] = -np.inf
        return lp
    np.random.seed(7)
    s = EnsembleSampler(10, 2, f, vectorize=True)
    s.run_mcmc(0.1 * np.random.randn(10, 2), 300, rstate0=np.random.get_state())
    assert np.all(s.chain[:, :, 0] <= 1.0) and np.all(np.isfinite(s.lnprobability))
    bad = EnsembleSampler(10, 2, lambda x: np.full(len(np.atleast_2d(x)), np.nan), vectorize=True)
    with pytest.raises(ValueError):
        bad.run_mcmc(np.zeros((10, 2)), 1)


def test_integrated_time_of_ar1_process():
    rng = np.random.default_rng(2)
    rho, n = 0.9, 20000
    x = np.zeros((n, 4, 1))
    for t in range(1, n):
        x[t] = rho * x[t - 1] + rng.standard_normal((4, 1))
    tau = integrated_time(x)[0]
    assert abs(tau - (1 + rho) / (1 - rho)) < 4.0          # 19 for rho = 0.9


def test_philox4x32_10_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (the counter RNG of the device-resident sampler)."""
    from lumfuncmcmc_b200.sampler import philox4x32_10
    r = philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(x[0]) for x in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    r = philox4x32_10([f], [f], [f], [f], f, f)
    assert [int(x[0]) for x in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = philox4x32_10([0x243f6a88], [0x85a308d3], [0x13198a2e], [0x03707344], 0xa4093822, 0x299f31d0)
    assert [int(x[0]) for x in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_stretch_reference_samples_a_gaussian():
    """The host restatement of the device sampler's algorithm is a valid sampler (fixed split, counter stream)."""
    from lumfuncmcmc_b200.sampler import philox_stretch_reference
    mu, sig = np.array([1.0, -2.0]), np.array([0.5, 2.0])
    rs = np.random.RandomState(2)
    chain, lnp, nacc = philox_stretch_reference(_gauss(mu, sig), mu + 0.1 * rs.randn(40, 2), 1500, seed=12345)
    flat = chain[300:].reshape(-1, 2)
    assert np.all(np.abs(flat.mean(axis=0) - mu) < 0.15 * sig)
    assert np.all(np.abs(flat.std(axis=0) / sig - 1.0) < 0.1)
    assert 0.3 < nacc.mean() / 1500 < 0.9
    # continuing a run (step0) reproduces the one-shot chain
    c1, l1, _ = philox_stretch_reference(_gauss(mu, sig), chain[9], 5, seed=12345, step0=10)
    assert np.array_equal(c1, chain[10:15])
