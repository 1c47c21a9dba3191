"""min_comp_frac > 0: per-source upper redshift limit and volume of the 1/V_eff estimator (SURVEY.md section 8 row a11,
reference lumfuncmcmc.py:521-524 -> VmaxLumFunc.py:739-753, :235-257).

CPU: the oracle's vectorised restatement is pinned (a) against the reference's own per-source fsolve + QUADPACK loop and
(b) against the weights the unmodified reference produced (golden ``veff_k2_mcf50``).  GPU: ``lf_veff_volumes`` against
both, through the C ABI, and against the restatement at 1e6 sources at 1e-12.

Tolerances against the reference's numbers: the upper redshift limit 1e-7 (fsolve's xtol is 1.5e-8; measured 6e-14), the
volume and the weights 1e-6: QUADPACK stops at its 50-subdivision limit on the kinked (piecewise-linear) interpolant and
reports an error estimate of 2e-7 of the integral itself (measured difference to the exact integral: <= 1.9e-7), so the
reference's volumes are only defined to that accuracy.  The device evaluates the integral of the interpolant exactly."""
import numpy as np
import pytest

from oracle import lf_oracle
from tests.test_host_setup import _build

MPC_CM = 3.085677581491367e24          # astropy's Mpc -> cm (VmaxLumFunc.py:737 ``.to('cm')``)


def _class_and_roots(n=250, nfields=2, seed=19):
    from lumfuncmcmc_b200 import configLF, synth
    from lumfuncmcmc_b200.lumfuncmcmc import LumFuncMCMC
    cat = synth.make_catalogue(n, seed=seed, nfields=nfields)
    m = LumFuncMCMC(cat['z'], flux=cat['flux'], flux_e=cat['flux_e'], Flim=list(cat['Flim']), alpha=cat['alpha'],
                    Omega_0=list(cat['Omega_0']), Flim_lims=configLF.Flim_lims, alpha_lims=configLF.alpha_lims,
                    sch_al=configLF.sch_al, Lstar=configLF.Lstar, phistar=configLF.phistar, fcmin=cat['fcmin'],
                    min_comp_frac=0.5, field_names=cat['field_names'], field_ind=cat['field_ind'], nbins=20, nboot=30)
    m.getFlim()
    return m, m.rootsf.ev(m.Flims_arr, m.alpha)          # per source, as the reference evaluates it (lumfuncmcmc.py:520)


def test_oracle_volumes_restatement_matches_the_reference_loop_and_golden(golden):
    from lumfuncmcmc_b200.cosmology import cosmo
    g = golden('veff_k2_mcf50')
    m, roots = _class_and_roots()
    zm_v, vol_v, ok_v = lf_oracle.veff_volumes_vectorised(m.lum, roots, m.zmin, m.zmax, m.dVdzf.x, m.dVdzf.y,
                                                          cosmo.luminosity_distance)
    zm_l, vol_l, ok_l = lf_oracle.veff_volumes_loop(m.lum, roots, m.zmin, m.zmax, m.dVdzf,
                                                    lambda z: cosmo.luminosity_distance(z) * MPC_CM)
    assert np.array_equal(ok_v, ok_l) and 0 < ok_v.sum() < len(ok_v)
    assert np.any(zm_v[ok_v] < m.zmax) and np.any(zm_v == m.zmax)        # both branches are populated
    np.testing.assert_allclose(zm_v[ok_v], zm_l[ok_v], rtol=1e-7)
    np.testing.assert_allclose(vol_v, vol_l, rtol=1e-6)
    # the weights the unmodified reference produced for this catalogue
    phi = lf_oracle.veff_weights(m.flux, m.Flims_arr, m.alpha, m.fcmin, sum(m.Omega_0), vol_v, m.zmin, zmaxval=zm_v)
    assert np.array_equal(phi == 0.0, g['phifunc'] == 0.0)
    nz = phi != 0
    np.testing.assert_allclose(phi[nz], g['phifunc'][nz], rtol=1e-6)


@pytest.mark.gpu
def test_device_volumes_match_the_reference_loop():
    from lumfuncmcmc_b200.cosmology import cosmo
    m, roots = _class_and_roots(n=1200, nfields=3, seed=23)
    fi = np.asarray(m.field_ind)
    zm_l, vol_l, ok_l = m._veff_volumes_host(np.array([roots[fi[k]] for k in range(m.nfields)]))   # the reference's loop, verbatim
    eng = m._veff_engine()
    eng.veff_set_sample(m.flux, m.lum, m.field_ind)
    eng.veff_set_volume_table(cosmo, m.dVdzf.x, m.dVdzf.y)
    fmin = np.array([roots[fi[k]] for k in range(m.nfields)])
    # the per-source spline evaluation of the reference equals the per-field one the class now does, bit for bit
    assert np.array_equal(np.repeat(m.rootsf.ev(np.asarray(m.Flim, dtype=float), np.full(m.nfields, float(m.alpha))), np.diff(fi)), roots)
    zm, vol, ok = eng.veff_volumes(m.zmin, m.zmax, float(cosmo.luminosity_distance(m.zmin)),
                                   float(cosmo.luminosity_distance(m.zmax)), fmin, want=True)
    assert np.array_equal(ok.astype(bool), ok_l.astype(bool)) and 0 < ok.sum() < len(ok)
    np.testing.assert_allclose(zm[ok_l > 0], zm_l[ok_l > 0], rtol=1e-7)
    np.testing.assert_allclose(vol, vol_l, rtol=1e-6)
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize('n', [1000000])
def test_device_volumes_large_against_restatement(n):
    """1e6 sources, the knot count the reference would use (one knot per source): Newton on the device D_L and the exact
    integral of the interpolant against the oracle's bisection + cumulative trapezoids; counts of the binned LF exact."""
    from lumfuncmcmc_b200 import synth
    from lumfuncmcmc_b200.cosmology import cosmo
    from lumfuncmcmc_b200.engine import VeffEngine
    cat = synth.make_catalogue(n, seed=77, nfields=5)
    z = np.concatenate(cat['z'])
    flux = 1.0e-17 * np.concatenate(cat['flux'])
    lum = np.log10(4.0 * np.pi * (cosmo.luminosity_distance(z) * 3.086e24) ** 2 * flux)
    fi = np.asarray(cat['field_ind'], dtype=np.int64)
    zmin, zmax = float(z.min()), float(z.max())
    zk = np.linspace(0.95 * zmin, 1.05 * zmax, n)
    dVk = cosmo.differential_comoving_volume(zk)
    fmin_field = 0.9e-17 * np.asarray(cat['Flim'])                      # a flux limit that cuts part of every field
    fmin_src = np.repeat(fmin_field, np.diff(fi))
    zm_o, vol_o, ok_o = lf_oracle.veff_volumes_vectorised(lum, fmin_src, zmin, zmax, zk, dVk, cosmo.luminosity_distance)
    eng = VeffEngine(device=0)
    eng.veff_set_sample(flux, lum, fi)
    eng.veff_set_volume_table(cosmo, zk, dVk)
    zm, vol, ok = eng.veff_volumes(zmin, zmax, float(cosmo.luminosity_distance(zmin)), float(cosmo.luminosity_distance(zmax)),
                                   fmin_field, want=True)
    # a source whose limit lies within rounding of zmin may fall on either side; nothing else may differ
    differ = ok.astype(bool) != ok_o
    assert differ.sum() <= 2 and np.all(np.abs(zm_o[differ] - zmin) < 1e-12)
    both = ok.astype(bool) & ok_o
    assert 0.05 * n < both.sum() < n and np.any(zm[both] < zmax) and np.any(zm[both] == zmax)
    np.testing.assert_allclose(zm[both], zm_o[both], rtol=1e-13)
    # volumes of sources whose limit is barely above zmin are differences of nearly equal integrals: absolute bound
    scale = vol_o[both].max()
    assert np.max(np.abs(vol[both] - vol_o[both])) <= 1e-12 * scale
    # weights + binned LF with the per-source volumes left on the device
    edges = np.linspace(lum.min() * 1.001, lum.max(), 51)
    phi, counts, sums = eng.veff_bin_resident(cat['Flim'], cat['alpha'], cat['fcmin'], float(np.sum(cat['Omega_0'])), 1.0, edges,
                                              device_volumes=True, want_phi=True)
    flims_arr = np.repeat(np.asarray(cat['Flim'], dtype=np.float64), np.diff(fi))
    ref = lf_oracle.veff_weights(flux, flims_arr, cat['alpha'], cat['fcmin'], float(np.sum(cat['Omega_0'])), vol_o, zmin,
                                 zmaxval=np.where(ok.astype(bool), zmax, zmin))
    big = both & (vol_o > 1e-6 * scale)
    np.testing.assert_allclose(phi[big], ref[big], rtol=1e-9)
    assert np.all(phi[~ok.astype(bool)] == 0.0)
    idx = np.searchsorted(edges, lum, side='right') - 1
    want = np.bincount(idx[(idx >= 0) & (idx < 50) & (lum < edges[-1])], minlength=50)
    assert np.array_equal(counts, want)
    eng.close()


@pytest.mark.gpu
def test_resident_sample_equals_host_buffer_path_and_weights_are_fetched_lazily():
    from lumfuncmcmc_b200 import synth
    from lumfuncmcmc_b200.cosmology import cosmo
    from lumfuncmcmc_b200.engine import VeffEngine
    n = 300000
    cat = synth.make_catalogue(n, seed=5, nfields=4)
    z = np.concatenate(cat['z'])
    flux = 1.0e-17 * np.concatenate(cat['flux'])
    lum = np.log10(4.0 * np.pi * (cosmo.luminosity_distance(z) * 3.086e24) ** 2 * flux)
    fi = np.asarray(cat['field_ind'], dtype=np.int64)
    edges = np.linspace(lum.min() * 1.001, lum.max(), 41)
    so = float(np.sum(cat['Omega_0']))
    a, b = VeffEngine(device=0), VeffEngine(device=0)
    phi_a, cnt_a, sum_a = a.veff_bin(flux, lum, fi, cat['Flim'], cat['alpha'], cat['fcmin'], so, 3.0e10, edges)
    b.veff_set_sample(flux, lum, fi)
    for flim, alpha in ((cat['Flim'], cat['alpha']), ([2.0, 3.0, 2.5, 4.0], 4.2), (cat['Flim'], cat['alpha'])):
        none, cnt_b, sum_b = b.veff_bin_resident(flim, alpha, cat['fcmin'], so, 3.0e10, edges)   # repeated calls, same sample
        assert none is None
    assert np.array_equal(cnt_a, cnt_b)
    np.testing.assert_allclose(sum_b, sum_a, rtol=1e-12)
    # the resident route forms n = alpha (u - log10(F50 / F0)) from the stored u = log10(f / F0) instead of taking the
    # logarithm of f / F50 on every call: same weights to the rounding of that logarithm
    np.testing.assert_allclose(b.veff_phi(), phi_a, rtol=1e-13)
    flims_arr = np.repeat(np.asarray(cat['Flim'], dtype=np.float64), np.diff(fi))
    ref = lf_oracle.veff_weights(flux, flims_arr, cat['alpha'], cat['fcmin'], so, 3.0e10, 0.0)
    np.testing.assert_allclose(b.veff_phi(), ref, rtol=1e-13)
    cnt_c, sum_c = b.bin_weights(None, None, edges)                               # re-binning the resident weights
    assert np.array_equal(cnt_c, cnt_a)
    np.testing.assert_allclose(sum_c, sum_a, rtol=1e-13)
    mult = np.bincount(np.random.RandomState(3).randint(n, size=n), minlength=n)
    ca, sa = a.boot_bin(mult)
    cb, sb = b.boot_bin(mult)
    assert np.array_equal(ca, cb)
    np.testing.assert_allclose(sb, sa, rtol=1e-12)
    # other edges, other number of bins: the cached rows and counts are rebuilt
    edges2 = np.linspace(lum.min() * 1.001, lum.max(), 26)
    _, cnt2, sum2 = b.veff_bin_resident(cat['Flim'], cat['alpha'], cat['fcmin'], so, 3.0e10, edges2)
    idx = np.searchsorted(edges2, lum, side='right') - 1
    keep = (idx >= 0) & (idx < 25)
    assert np.array_equal(cnt2, np.bincount(idx[keep], minlength=25)[:25])
    np.testing.assert_allclose(sum2, np.bincount(idx[keep], weights=ref[keep], minlength=25)[:25], rtol=1e-12)
    for nb3 in (100, 200):                        # more bins than the 16-column kernel holds / than either resident kernel holds
        edges3 = np.linspace(lum.min() * 1.001, lum.max(), nb3 + 1)
        _, cnt4, sum4 = b.veff_bin_resident(cat['Flim'], cat['alpha'], cat['fcmin'], so, 3.0e10, edges3)
        idx3 = np.searchsorted(edges3, lum, side='right') - 1
        keep3 = (idx3 >= 0) & (idx3 < nb3)
        assert np.array_equal(cnt4, np.bincount(idx3[keep3], minlength=nb3)[:nb3])
        np.testing.assert_allclose(sum4, np.bincount(idx3[keep3], weights=ref[keep3], minlength=nb3)[:nb3], rtol=1e-12)
    with pytest.raises(Exception, match='positive'):
        b.veff_bin_resident([2.0, -1.0, 2.5, 4.0], cat['alpha'], cat['fcmin'], so, 3.0e10, edges2)
    _, cnt3, sum3 = b.veff_bin_resident(cat['Flim'], cat['alpha'], 0.0, so, 3.0e10, edges2)        # plain Fleming curve
    ref3 = lf_oracle.veff_weights(flux, flims_arr, cat['alpha'], 0.0, so, 3.0e10, 0.0)
    np.testing.assert_allclose(b.veff_phi(), ref3, rtol=1e-13)
    assert np.array_equal(cnt3, cnt2)
    a.close()
    b.close()


@pytest.mark.gpu
def test_class_veff_minimum_completeness_runs_on_the_device(golden):
    """The class path for min_comp_frac > 0 no longer loops over sources on the host."""
    g = golden('veff_k2_mcf50')
    m, _ = _class_and_roots()
    called = []
    m._veff_volumes_host = lambda *a, **k: called.append(1)
    np.random.seed(int(g['seed']))
    m.VeffLF()
    assert not called
    assert m._phifunc is None                                   # still on the device ...
    phi = m.phifunc                                             # ... until somebody reads the attribute
    assert m._phifunc is phi and phi.shape == m.lum.shape
    assert np.array_equal(phi == 0.0, g['phifunc'] == 0.0)
    nz = g['phifunc'] != 0
    np.testing.assert_allclose(phi[nz], g['phifunc'][nz], rtol=1e-6)
    assert np.array_equal(m.bincounts, g['counts'])
    np.testing.assert_allclose(m.lfbinorig, g['lfbinorig'], rtol=1e-6)
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize('n', [400, 65537, 1000000])
def test_device_mt19937_bootstrap_is_numpys_stream(n):
    """Replicates drawn on the device from NumPy's legacy MT19937 stream (reference VmaxLumFunc.py:353) are bit-identical to
    the host-drawn ones, and the host generator ends where the reference's would."""
    from lumfuncmcmc_b200 import VmaxLumFunc as V
    from lumfuncmcmc_b200.engine import VeffEngine
    rs = np.random.RandomState(n)
    L = rs.uniform(41.0, 44.0, n)
    phi = 10 ** rs.uniform(-6, -2, n)
    eng = VeffEngine(device=0)
    out, after, nxt = {}, {}, {}
    for rng in ('host', 'mt19937'):
        np.random.seed(20260 + n)
        np.random.rand(int(n) % 700 + 3)                          # start somewhere inside a state block
        out[rng] = V.getBootErrLog(L, phi, 1.2, 1.9, nboot=7, nbin=12, engine=eng, return_counts=True, rng=rng)
        after[rng] = np.random.get_state()
        nxt[rng] = np.random.randint(10 ** 6, size=8)
    for a, b in zip(out['host'], out['mt19937']):
        assert np.array_equal(a, b)
    assert after['host'][2] == after['mt19937'][2] and np.array_equal(after['host'][1], after['mt19937'][1])
    assert np.array_equal(nxt['host'], nxt['mt19937'])
    eng.close()


@pytest.mark.gpu
def test_golden_bootstrap_variance_with_the_device_stream(golden):
    """veff_k3_n400: the variances the unmodified reference produced, from the stream generated on the device."""
    from lumfuncmcmc_b200 import VmaxLumFunc as V
    from lumfuncmcmc_b200.engine import VeffEngine
    g = golden('veff_k3_n400')
    eng = VeffEngine(device=0)
    np.random.seed(int(g['seed']))
    Lavg, lfb, var = V.getBootErrLog(g['lum'], g['phifunc'], float(g['zmin']), float(g['zmax']), nboot=int(g['nboot']),
                                     nbin=int(g['nbins']), engine=eng, rng='mt19937')
    np.testing.assert_array_equal(Lavg, g['Lavg'])
    np.testing.assert_allclose(lfb, g['lfbinorig'], rtol=1e-12)
    np.testing.assert_allclose(var, g['var'], rtol=1e-9)
    eng.close()


@pytest.mark.gpu
def test_per_source_omega_on_the_device_matches_numpy():
    from lumfuncmcmc_b200 import synth
    from lumfuncmcmc_b200.cosmology import cosmo
    from lumfuncmcmc_b200.lfbase import Omega
    from lumfuncmcmc_b200.setup_gpu import LinearTable, omega_sources
    n = 250000
    cat = synth.make_catalogue(n, seed=9, nfields=4)
    z = np.concatenate(cat['z'])
    flux = 1.0e-17 * np.concatenate(cat['flux'])
    lum = np.log10(4.0 * np.pi * (cosmo.luminosity_distance(z) * 3.086e24) ** 2 * flux)
    fi = np.asarray(cat['field_ind'], dtype=np.int64)
    zint = np.linspace(0.95 * z.min(), 1.05 * z.max(), 5000)
    DLf = LinearTable(zint, cosmo.luminosity_distance(zint))
    om0 = np.asarray(cat['Omega_0']).astype(int)
    for fcmin in (cat['fcmin'], 0.0):
        got = omega_sources(lum, z, fi, om0, cat['Flim'], cat['alpha'], fcmin, DLf)
        want = Omega(lum, z, lambda zz: np.interp(zz, zint, DLf.y), np.repeat(om0, np.diff(fi)),
                     1.0e-17 * np.repeat(np.asarray(cat['Flim'], dtype=np.float64), np.diff(fi)), cat['alpha'], fcmin)
        np.testing.assert_allclose(got, want, rtol=2e-14)
