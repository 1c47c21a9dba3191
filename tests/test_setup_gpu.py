"""GPU: set-up tables on the device against the host arithmetic they replace."""
import numpy as np
import pytest

from lumfuncmcmc_b200 import setup_gpu
from lumfuncmcmc_b200.cosmology import LambdaCDM, cosmo

pytestmark = pytest.mark.gpu


def test_interp_linear_is_numpy_interp_bit_for_bit():
    rng = np.random.default_rng(0)
    for nk, uniform in ((300000, True), (5000, False), (2, True)):
        xk = np.linspace(1.102, 1.995, nk) if uniform else np.sort(rng.uniform(1.1, 2.0, nk)) ** 3
        yk = np.sqrt(xk) * 1234.5 + rng.normal(size=nk)
        x = np.concatenate([rng.uniform(xk[0], xk[-1], 400000), xk[:50], [xk[-1], xk[0]]])
        got = setup_gpu.interp_linear(xk, yk, x)
        assert np.array_equal(got, np.interp(x, xk, yk))
    with pytest.raises(ValueError):
        setup_gpu.interp_linear(xk, yk, np.array([xk[-1] + 1.0]))
    with pytest.raises(ValueError):
        setup_gpu.interp_linear(xk, yk, np.array([np.nan]))
    # the drop-in table object: scalar and array calls, GPU for long arrays, host for short ones -- same numbers
    t = setup_gpu.LinearTable(xk, yk)
    long_x = rng.uniform(xk[0], xk[-1], setup_gpu.GPU_MIN_POINTS + 7)
    assert np.array_equal(t(long_x), np.interp(long_x, xk, yk))
    assert t(xk[0]) == yk[0] and np.array_equal(t(long_x[:9].reshape(3, 3)), np.interp(long_x[:9], xk, yk).reshape(3, 3))


@pytest.mark.parametrize('c', [cosmo, LambdaCDM(H0=67.7, Om0=0.31, Ode0=0.75, Tcmb0=0.0), LambdaCDM(H0=70, Om0=0.3, Ode0=0.6, Tcmb0=2.725)])
def test_cosmology_distances_match_the_host(c):
    """Closed (the reference's: Ok0 = -Or0 < 0), flat-ish and open curvature branches; 1e-15 relative (sin/sinh ulps)."""
    z = np.concatenate([np.random.default_rng(1).uniform(0.0, 3.0, 300000), [0.0, 1.0e-9, 2.999999]])
    dl, dv = setup_gpu.cosmo_distances(c, z)
    np.testing.assert_allclose(dl, c.luminosity_distance(z), rtol=2e-15, atol=0)
    np.testing.assert_allclose(dv, c.differential_comoving_volume(z), rtol=4e-15, atol=0)
    with pytest.raises(Exception):
        setup_gpu.cosmo_distances(c, np.array([-0.1]))


def test_large_catalogue_setup_uses_the_gpu_and_equals_the_host_tables(monkeypatch):
    """Constructor chain at N above the GPU threshold: tables from the device path equal the host path to 1e-15 and
    the per-source interpolations are bit-identical; lnprob of the two set-ups agrees to 1e-12."""
    from lumfuncmcmc_b200 import configLF, synth
    from lumfuncmcmc_b200.lumfuncmcmc import LumFuncMCMC
    n = setup_gpu.GPU_MIN_POINTS + 50000
    cat = synth.make_catalogue(n, seed=41, nfields=3)

    def build():
        return LumFuncMCMC(cat['z'], flux=cat['flux'], flux_e=cat['flux_e'], Flim=list(cat['Flim']), alpha=cat['alpha'],
                           Omega_0=list(cat['Omega_0']), Flim_lims=configLF.Flim_lims, alpha_lims=configLF.alpha_lims,
                           sch_al=configLF.sch_al, Lstar=configLF.Lstar, phistar=configLF.phistar, fcmin=cat['fcmin'],
                           min_comp_frac=0.0, field_names=cat['field_names'], field_ind=cat['field_ind'])
    calls = {'cosmo': 0, 'interp': 0, 'omega': 0}
    real_c, real_i, real_o = setup_gpu.cosmo_distances, setup_gpu.interp_linear, setup_gpu.omega_sources

    def count_c(*a, **k):
        calls['cosmo'] += 1
        return real_c(*a, **k)

    def count_i(*a, **k):
        calls['interp'] += 1
        return real_i(*a, **k)
    def count_o(*a, **k):
        calls['omega'] += 1
        return real_o(*a, **k)
    import lumfuncmcmc_b200.lfbase as lfbase
    monkeypatch.setattr(lfbase, 'cosmo_distances', count_c)
    monkeypatch.setattr(setup_gpu, 'interp_linear', count_i)
    monkeypatch.setattr(setup_gpu, 'omega_sources', count_o)
    g = build()
    # the three O(N) passes of the constructor: D_L per source, the N-knot tables, the per-source tabulated Omega (which
    # interpolates D_L inside the kernel)
    assert calls['cosmo'] == 2 and calls['omega'] == 1
    monkeypatch.setattr(lfbase, 'gpu_count', lambda: 0)
    monkeypatch.setattr(setup_gpu, 'gpu_count', lambda: 0)
    h = build()
    np.testing.assert_allclose(g.DLf.y, h.DLf.y, rtol=2e-15, atol=0)
    np.testing.assert_allclose(g.dVdzf.y, h.dVdzf.y, rtol=4e-15, atol=0)
    np.testing.assert_allclose(g.lum, h.lum, rtol=1e-15, atol=0)
    # device set-up path: same formulas and order of operations with libdevice sin / sqrt / pow / log10 / exp instead of
    # NumPy's -- pinned here at 2e-15 (distances), 4e-15 (dV/dz), 1e-15 (lum) and 1e-13 (Omega per source)
    np.testing.assert_allclose(g.Om_arr, h.Om_arr, rtol=1e-13, atol=0)
    th = synth.draw_thetas(g.engine_inputs(), 'free', 8, seed=3, mode='near', scale=0.02)
    a, b = g.lnprob(th), h.lnprob(th)
    assert np.max(np.abs(a - b) / np.abs(b)) < 1e-12
    g.close()
    h.close()
