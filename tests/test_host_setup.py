"""CPU: the drop-in classes' set-up chain reproduces, bit for bit, the arrays the reference's own constructors
produced for the same catalogue (stored in the golden fixtures by oracle/make_golden.py)."""
import numpy as np
import pytest

from lumfuncmcmc_b200 import configLF, synth

KEYS = ['lum', 'z', 'zint', 'DLarr', 'dVdzarr', 'logL', 'zarr', 'DL_zarr', 'volume_part', 'Om_arr', 'integ_part']


def _build(kind, n, nfields, seed, mcf=0.0, fix_sch_al=False, evolve=None):
    cat = synth.make_catalogue(n, seed=seed, nfields=nfields, evolve=evolve)
    np.random.seed(seed)
    common = dict(flux=cat['flux'], flux_e=cat['flux_e'], Flim=list(cat['Flim']), alpha=cat['alpha'],
                  Omega_0=list(cat['Omega_0']), sch_al=configLF.sch_al, sch_al_lims=configLF.sch_al_lims,
                  Lstar=configLF.Lstar, Lstar_lims=configLF.Lstar_lims, phistar=configLF.phistar,
                  phistar_lims=configLF.phistar_lims, Lc=configLF.Lc, Lh=configLF.Lh, fcmin=cat['fcmin'],
                  min_comp_frac=mcf, field_names=cat['field_names'], field_ind=cat['field_ind'], fix_sch_al=fix_sch_al)
    if kind == 'z':
        from lumfuncmcmc_b200.lumfuncmcmc_z import LumFuncMCMCz
        return LumFuncMCMCz(cat['z'], z1=1.20, z2=1.53, z3=1.86, **common)
    from lumfuncmcmc_b200.lumfuncmcmc import LumFuncMCMC
    return LumFuncMCMC(cat['z'], Flim_lims=configLF.Flim_lims, alpha_lims=configLF.alpha_lims,
                       fix_comp=(kind == 'fixed'), **common)


@pytest.mark.parametrize('name,kind,args', [
    ('free_k5_n2000', 'free', dict(n=2000, nfields=5, seed=11)),
    ('fixed_k2_n800', 'fixed', dict(n=800, nfields=2, seed=14)),
    ('z_k2_n800', 'z', dict(n=800, nfields=2, seed=16, evolve=(0.3, -0.2))),
    ('free_k2_mcf50', 'free', dict(n=600, nfields=2, seed=13, mcf=0.5)),
])
def test_setup_tables_match_reference(golden, name, kind, args):
    g = golden(name)
    m = _build(kind, **args)
    inp = m.engine_inputs()
    for key in KEYS:
        assert np.array_equal(np.asarray(inp[key]), g[key]), key
    assert np.array_equal(inp['field_ind'], g['field_ind'])
    assert np.array_equal(m.flux, g['flux']) and np.array_equal(m.lum_e, g['lum_e'])
    assert m.Omega_0_arr.dtype.kind == 'i'
    assert m.size_ln == len(g['zarr'])
    # every logL[k] is one aliased grid
    assert all(m.logL[k] is m.logL[0] for k in range(m.nfields))


def test_parameter_bookkeeping_matches_reference_layout():
    m = _build('free', n=300, nfields=3, seed=3)
    assert m.get_param_names()[:3] == [r'$\log L_*$', r'$\log \phi_*$', r'$\alpha$']
    assert len(m.get_param_names()) == 3 + 3 + 1 == len(m.get_params())
    np.random.seed(1)
    pos = m.get_init_walker_values()
    assert pos.shape == (100, 7)
    np.random.seed(1)
    u = np.random.rand(100, 7)
    lims = np.array([m.Lstar_lims, m.phistar_lims, m.sch_al_lims] + [m.Flim_lims] * 3 + [m.alpha_lims], dtype=float)
    assert np.array_equal(pos, u * (lims[:, 1] - lims[:, 0]) + lims[:, 0])
    m.set_parameters_from_list(pos[0])
    assert m.Lstar == pos[0, 0] and m.sch_al == pos[0, 2] and m.alpha == pos[0, 6]
    assert m.lnprior() == 0.0
    m.alpha = 99.0
    assert m.lnprior() == -np.inf
    th = m._theta_matrix(pos[:4], True)
    assert th.shape == (4, 7) and np.array_equal(th, pos[:4])


def test_lnprob_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from lumfuncmcmc_b200._lib import EngineError
    m = _build('free', n=200, nfields=2, seed=4)
    with pytest.raises(EngineError):
        m.lnprob(np.array([42.5, -2.0, -1.49, 2.72, 3.61, 4.56]))


def test_linear_table_host_path_is_interp1d():
    """Without a GPU (or below the size threshold) LinearTable evaluates numpy.interp = interp1d(kind='linear')."""
    from scipy.interpolate import interp1d
    from lumfuncmcmc_b200.setup_gpu import LinearTable
    rng = np.random.default_rng(5)
    xk = np.linspace(1.1, 2.0, 5001)
    yk = np.sqrt(xk) * 321.0 + rng.normal(size=xk.size)
    t, ref = LinearTable(xk, yk), interp1d(xk, yk)
    x = np.concatenate([rng.uniform(1.1, 2.0, 20000), xk[:7], [xk[-1]]])
    assert np.array_equal(t(x), ref(x))
    assert t(1.5) == float(ref(1.5)) and np.array_equal(t(x[:6].reshape(2, 3)), ref(x[:6].reshape(2, 3)))
    assert np.array_equal(t.x, xk) and np.array_equal(t.y, yk)
    import pytest
    with pytest.raises(ValueError):
        t(np.array([2.5]))
    with pytest.raises(ValueError):
        LinearTable(xk[:1], yk[:1])
