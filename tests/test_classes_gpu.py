"""GPU: the drop-in classes end to end (constructor -> engine -> lnprob / fit_model / VeffLF) against the reference's
golden values and the oracle."""
import numpy as np
import pytest

from oracle import lf_oracle
from tests.test_host_setup import _build

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _check(got, ref):
    assert np.array_equal(np.isneginf(got), np.isneginf(ref))
    fin = np.isfinite(ref)
    assert np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])) <= RTOL


@pytest.mark.parametrize('name,kind,args', [
    ('free_k5_n2000', 'free', dict(n=2000, nfields=5, seed=11)),
    ('free_k3_fixal', 'free', dict(n=900, nfields=3, seed=12, fix_sch_al=True)),
    ('fixed_k2_n800', 'fixed', dict(n=800, nfields=2, seed=14)),
    ('fixed_k2_fixal', 'fixed', dict(n=500, nfields=2, seed=15, fix_sch_al=True)),
    ('z_k2_n800', 'z', dict(n=800, nfields=2, seed=16, evolve=(0.3, -0.2))),
    ('z_k2_fixal', 'z', dict(n=500, nfields=2, seed=17, fix_sch_al=True)),
])
def test_class_lnprob_matches_reference(golden, name, kind, args):
    g = golden(name)
    m = _build(kind, **args)
    fn = m.lnprob_fix_comp if kind == 'fixed' else m.lnprob
    th, ref = g['thetas'], g['lnprob_ref']
    _check(fn(th), ref)                                           # whole ensemble, one call
    for i in (0, 3, len(th) - 1):                                 # scalar API: float, and self is updated
        v = fn(th[i])
        assert isinstance(v, float)
        assert (np.isneginf(v) and np.isneginf(ref[i])) or abs(v - ref[i]) <= RTOL * abs(ref[i])
        if kind == 'z':
            assert np.array_equal([m.L1, m.phi3], [th[i][0], th[i][5]], equal_nan=True)
        else:
            assert np.array_equal([m.Lstar, m.phistar], [th[i][0], th[i][1]], equal_nan=True)
    # lnlike() evaluates at the current attributes and ignores the prior
    m.set_parameters_from_list(th[0])
    model = lf_oracle.make_model(g, kind)
    model.unpack(th[0])
    want = model.lnlike_fix_comp() if kind == 'fixed' else model.lnlike()
    got = m.lnlike_fix_comp() if kind == 'fixed' else m.lnlike()
    assert abs(got - want) <= RTOL * abs(want)
    m.close()


def test_veff_lf_matches_reference(golden):
    g = golden('veff_k3_n400')
    from lumfuncmcmc_b200 import configLF, synth
    from lumfuncmcmc_b200.lumfuncmcmc import LumFuncMCMC
    cat = synth.make_catalogue(400, seed=18, nfields=3)
    m = LumFuncMCMC(cat['z'], flux=cat['flux'], flux_e=cat['flux_e'], Flim=list(cat['Flim']), alpha=cat['alpha'],
                    Omega_0=list(cat['Omega_0']), Flim_lims=configLF.Flim_lims, alpha_lims=configLF.alpha_lims,
                    sch_al=configLF.sch_al, Lstar=configLF.Lstar, phistar=configLF.phistar, fcmin=cat['fcmin'],
                    min_comp_frac=0.0, field_names=cat['field_names'], field_ind=cat['field_ind'], nbins=20, nboot=30)
    np.random.seed(int(g['seed']))
    m.VeffLF()
    np.testing.assert_allclose(m.phifunc, g['phifunc'], rtol=1e-13)
    assert np.array_equal(m.bincounts, g['counts'])
    np.testing.assert_array_equal(m.Lavg, g['Lavg'])
    np.testing.assert_allclose(m.lfbinorig, g['lfbinorig'], rtol=1e-12)
    np.testing.assert_allclose(m.var, g['var'], rtol=1e-9)         # same RNG stream -> same resamples
    m.close()


def test_veff_lf_with_minimum_completeness(golden):
    """min_comp_frac > 0: per-source upper redshift limit and volume (reference lumfuncmcmc.py:522-524)."""
    g = golden('veff_k2_mcf50')
    from lumfuncmcmc_b200 import configLF, synth
    from lumfuncmcmc_b200.lumfuncmcmc import LumFuncMCMC
    cat = synth.make_catalogue(250, seed=19, nfields=2)
    m = LumFuncMCMC(cat['z'], flux=cat['flux'], flux_e=cat['flux_e'], Flim=list(cat['Flim']), alpha=cat['alpha'],
                    Omega_0=list(cat['Omega_0']), Flim_lims=configLF.Flim_lims, alpha_lims=configLF.alpha_lims,
                    sch_al=configLF.sch_al, Lstar=configLF.Lstar, phistar=configLF.phistar, fcmin=cat['fcmin'],
                    min_comp_frac=0.5, field_names=cat['field_names'], field_ind=cat['field_ind'], nbins=20, nboot=30)
    np.random.seed(int(g['seed']))
    m.VeffLF()
    assert np.array_equal(m.phifunc == 0.0, g['phifunc'] == 0.0)
    nz = g['phifunc'] != 0
    # the reference's per-source QUADPACK volumes carry their own error estimate of 2e-7 (50-subdivision limit on the
    # kinked interpolant); the device integrates the interpolant exactly (tests/test_veff_volumes.py)
    np.testing.assert_allclose(m.phifunc[nz], g['phifunc'][nz], rtol=1e-6)
    assert np.array_equal(m.bincounts, g['counts'])
    np.testing.assert_allclose(m.lfbinorig, g['lfbinorig'], rtol=1e-6)
    m.close()


def test_fit_model_small_run_and_chain_replay():
    """Same sampler code and seed driven by the engine and by the oracle: identical accept/reject decisions, so
    identical chains; posterior summaries follow."""
    from lumfuncmcmc_b200.sampler import EnsembleSampler
    m = _build('free', n=600, nfields=2, seed=31)
    m.nwalkers, m.nsteps = 16, 25
    inp = m.engine_inputs()
    np.random.seed(77)
    truth = np.array([42.5, -3.2, -1.49] + list(m.Flim) + [m.alpha])
    pos = truth + 0.01 * np.random.randn(m.nwalkers, len(truth))
    state = np.random.get_state()
    eng = EnsembleSampler(m.nwalkers, len(truth), m.lnprob, vectorize=True)
    eng.run_mcmc(pos, m.nsteps, rstate0=state)
    ora = EnsembleSampler(m.nwalkers, len(truth), lambda th: lf_oracle.lnprob_batch(inp, 'free', th), vectorize=True)
    ora.run_mcmc(pos, m.nsteps, rstate0=state)
    assert np.array_equal(eng.chain, ora.chain)
    np.testing.assert_allclose(eng.lnprobability, ora.lnprobability, rtol=RTOL)
    # the class's own driver
    np.random.seed(78)
    m.fit_model()
    assert m.samples.shape[1] == len(truth) + 1 and np.isfinite(m.samples[:, -1]).all()
    assert m.chain.shape == (16, 25, len(truth))
    from lumfuncmcmc_b200.tableio import Table
    names = m.get_param_names()
    labels = ['Line'] + [n + '_%02d' % p for n in names for p in (16, 50, 84)]
    m.table = Table(names=labels, dtype=['S10'] + ['f8'] * (len(labels) - 1))
    m.table.add_row(['OIII'] + [0.] * (len(labels) - 1))
    m.add_fitinfo_to_table([16, 50, 84])
    med = np.array([m.table[-1][2 + 3 * j] for j in range(len(names))])
    assert np.all(np.isfinite(med)) and abs(med[0] - 42.5) < 1.0
    m.nboot, m.nbins = 5, 10
    m.set_median_fit(rndsamples=20)
    assert m.medianLF.shape == m.lum.shape and len(m.Lavg) == 10 and np.all(m.var > 0)
    m.close()


@pytest.mark.parametrize('evolving', [False, True])
def test_driver_end_to_end_writes_the_reference_outputs(tmp_path, monkeypatch, evolving):
    """Config-1 in miniature through the unchanged command line: catalogue file in, fitposterior / bestfitLF / VeffLF /
    parameter table / .args out (reference run_lumfuncmcmc.py:297-330)."""
    import os
    from lumfuncmcmc_b200 import synth
    from lumfuncmcmc_b200.tableio import Table
    from tests.test_driver_cpu import _write_catalogue
    cat = synth.make_catalogue(1500, seed=51, evolve=(0.3, -0.2) if evolving else None)
    _write_catalogue(str(tmp_path / 'cat.dat'), cat)
    monkeypatch.chdir(tmp_path)
    np.random.seed(5)
    if evolving:
        import run_lumfuncmcmc_z as drv
        out, nd = 'LFMCMCzOut', 7
    else:
        import run_lumfuncmcmc as drv
        out, nd = 'LFMCMCOut', 9
    m = drv.main(['-f', 'cat.dat', '-o', 'fit.dat', '-nw', '20', '-ns', '30', '-nboot', '5', '-nbins', '12'])
    tag = 'fit_nb12_nw20_ns30_mcf0'
    for stem in ('fitposterior_%s.dat' % tag, 'bestfitLF_%s.dat' % tag, 'VeffLF_%s.dat' % tag, 'fit.dat', 'fit.dat.args'):
        assert os.path.isfile(os.path.join(out, stem)), stem
    post = Table.read(os.path.join(out, 'fitposterior_%s.dat' % tag))
    assert post.colnames[-1] == 'Ln Prob' and len(post.colnames) == nd + 1
    assert np.array_equal(post.as_array(), m.samples)
    veff = Table.read(os.path.join(out, 'VeffLF_%s.dat' % tag))
    assert len(veff) == 12 and np.all(veff['BinLFErr'] > 0)
    # second invocation hits the result cache and only re-summarises
    m2 = drv.main(['-f', 'cat.dat', '-o', 'fit.dat', '-nw', '20', '-ns', '30', '-nboot', '5', '-nbins', '12'])
    assert np.array_equal(m2.samples, m.samples) and not hasattr(m2, 'chain')
    m.close()
    m2.close()


@pytest.mark.parametrize('kind', ['free', 'z'])
def test_device_sampler_chain_equals_host_replay(kind):
    """The device-resident sampler (Philox stream, CUDA graph per update) produces, sample for sample, the chain of
    its host restatement driven by the same engine's lnprob; stored ln-probabilities are the engine's values at the
    stored positions and agree with the oracle to 1e-10."""
    from lumfuncmcmc_b200 import synth
    from lumfuncmcmc_b200.engine import LikelihoodEngine
    from lumfuncmcmc_b200.sampler import DeviceEnsembleSampler, philox_stretch_reference
    cat = synth.make_catalogue(3000, seed=31, nfields=3, evolve=(0.3, -0.2) if kind == 'z' else None)
    inp = synth.direct_inputs(cat, nknots=512, size_ln=101 if kind == 'free' else 201, tabulated=(kind != 'free'))
    eng = LikelihoodEngine(inp, kind, device=0)
    W, nsteps, seed = 48, 25, 0x1234567890abcdef
    p0 = np.concatenate([synth.draw_thetas(inp, kind, W - 6, seed=2, mode='near', scale=0.01),
                         synth.draw_thetas(inp, kind, 6, seed=3, mode='prior')])         # a few walkers start at -inf
    smp = DeviceEnsembleSampler(W, eng.ndim, eng, seed=seed)
    pos, lp, _ = smp.run_mcmc(p0, nsteps)
    chain, lnp, nacc = philox_stretch_reference(eng.lnprob, p0, nsteps, seed)
    assert np.array_equal(smp.chain, np.swapaxes(chain, 0, 1))
    assert np.array_equal(smp.lnprobability, lnp.T)
    assert np.array_equal(smp.naccepted, nacc) and nacc.sum() > 0
    assert np.array_equal(pos, chain[-1]) and np.array_equal(lp, lnp[-1])
    # continuing the run continues the counter stream
    pos2, lp2, _ = smp.run_mcmc(pos, 5)
    chain2, lnp2, _ = philox_stretch_reference(eng.lnprob, chain[-1], 5, seed, step0=nsteps)
    assert np.array_equal(pos2, chain2[-1]) and smp.chain.shape == (W, nsteps + 5, eng.ndim)
    # the stored ln-probabilities are the oracle's values at the stored positions
    ref = lf_oracle.lnprob_batch(inp, kind, chain[-1])
    _check(lnp[-1], ref)
    assert smp.device_ms > 0.0
    eng.close()


def test_fit_model_on_the_device_sampler():
    m = _build('free', n=2000, nfields=5, seed=11)
    m.nwalkers, m.nsteps = 40, 30
    m.sampler_backend = 'device'
    np.random.seed(4)
    m.fit_model()
    from lumfuncmcmc_b200.sampler import DeviceEnsembleSampler
    assert isinstance(m.sampler, DeviceEnsembleSampler)
    assert m.chain.shape == (40, 30, 9) and m.samples.shape[1] == 10
    # every stored sample carries the engine's own lnprob of that position
    last = m.chain[:, -1, :]
    assert np.allclose(m.sampler.lnprobability[:, -1], m.lnprob(last), rtol=1e-12, atol=0, equal_nan=True)
    m.close()
