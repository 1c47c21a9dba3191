"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):   python oracle/make_golden.py

For every case it (1) builds a synthetic catalogue, (2) constructs the reference's own ``LumFuncMCMC`` /
``LumFuncMCMCz`` (imported through ``oracle/refstub.py``), (3) dumps the arrays the likelihood reads from
``self`` ("engine inputs"), (4) evaluates the reference's ``lnprob`` on a mixed bag of walker positions
(uniform-prior draws as in the reference's initialisation, near-truth draws, out-of-prior / NaN / boundary
rows), and (5) asserts that ``oracle/lf_oracle.py`` reproduces the reference bit-for-bit before writing.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refstub, lf_oracle            # noqa: E402
from lumfuncmcmc_b200 import synth, configLF     # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def inputs_from_reference(m, kind):
    """Arrays the reference's lnlike reads from ``self`` (lumfuncmcmc.py:360-393, lumfuncmcmc_z.py:364-376)."""
    inp = dict(lum=m.lum, z=m.z, zint=m.DLf.x, DLarr=m.DLf.y, dVdzarr=m.dVdzf.y,
               field_ind=np.asarray(m.field_ind, dtype=np.int64), Omega_0=np.asarray(m.Omega_0, dtype=np.float64),
               Flim=np.asarray(m.Flim, dtype=np.float64), alpha=float(m.alpha), fcmin=float(m.fcmin),
               logL=np.stack(m.logL), zarr=m.zarr, DL_zarr=m.DL_zarr, volume_part=m.volume_part,
               Om_arr=m.Om_arr, integ_part=np.stack(m.integ_part), flux=m.flux, lum_e=m.lum_e,
               zmin=float(m.zmin), zmax=float(m.zmax), min_comp_frac=float(m.min_comp_frac),
               Lstar_lims=np.asarray(m.Lstar_lims, dtype=np.float64),
               phistar_lims=np.asarray(m.phistar_lims, dtype=np.float64),
               sch_al_lims=np.asarray(m.sch_al_lims, dtype=np.float64),
               sch_al=float(m.sch_al), fix_sch_al=bool(m.fix_sch_al), Lc=float(m.Lc), Lh=float(m.Lh))
    if kind == 'z':
        inp.update(z1=float(m.z1), z2=float(m.z2), z3=float(m.z3))
    else:
        inp.update(Flim_lims=np.asarray(m.Flim_lims, dtype=np.float64),
                   alpha_lims=np.asarray(m.alpha_lims, dtype=np.float64), fix_comp=bool(m.fix_comp))
    return inp


def special_rows(thetas, lims):
    """Rows exercising the prior gate: one parameter out of the box at a time, exact bounds, NaN."""
    base = thetas[0].copy()
    rows = []
    for d in range(len(base)):
        for val in (lims[d][0] - 1e-9, lims[d][1] + 1e-9, lims[d][0], lims[d][1]):
            r = base.copy()
            r[d] = val
            rows.append(r)
    r = base.copy()
    r[0] = np.nan
    rows.append(r)
    r = base.copy()
    r[-1] = np.nan
    rows.append(r)
    return np.array(rows)


def param_lims(inp, kind):
    K = len(inp['Flim'])
    if kind == 'z':
        lims = [inp['Lstar_lims']] * 3 + [inp['phistar_lims']] * 3
        if not inp['fix_sch_al']:
            lims = lims + [inp['sch_al_lims']]
        return lims
    lims = [inp['Lstar_lims'], inp['phistar_lims']]
    if not inp['fix_sch_al']:
        lims = lims + [inp['sch_al_lims']]
    if kind == 'free':
        lims = lims + [inp['Flim_lims']] * K + [inp['alpha_lims']]
    return lims


def make_case(name, kind, n, nfields, seed, mcf=0.0, fix_sch_al=False, nprior=24, nnear=24, evolve=None):
    V, lf, lfz = refstub.load_reference()
    cat = synth.make_catalogue(n, seed=seed, nfields=nfields, evolve=evolve)
    np.random.seed(seed)          # LumFuncMCMCz draws L1..3, phi1..3 in its constructor (lumfuncmcmc_z.py:206-207)
    common = dict(flux=cat['flux'], flux_e=cat['flux_e'], Flim=list(cat['Flim']), alpha=cat['alpha'],
                  Omega_0=list(cat['Omega_0']), sch_al=configLF.sch_al, sch_al_lims=configLF.sch_al_lims,
                  Lstar=configLF.Lstar, Lstar_lims=configLF.Lstar_lims, phistar=configLF.phistar,
                  phistar_lims=configLF.phistar_lims, Lc=configLF.Lc, Lh=configLF.Lh, fcmin=cat['fcmin'],
                  min_comp_frac=mcf, field_names=cat['field_names'], field_ind=cat['field_ind'],
                  fix_sch_al=fix_sch_al)
    if kind == 'z':
        m = lfz.LumFuncMCMCz(cat['z'], z1=1.20, z2=1.53, z3=1.86, **common)
        fn = m.lnprob
    else:
        m = lf.LumFuncMCMC(cat['z'], Flim_lims=configLF.Flim_lims, alpha_lims=configLF.alpha_lims,
                           fix_comp=(kind == 'fixed'), **common)
        fn = m.lnprob_fix_comp if kind == 'fixed' else m.lnprob
    inp = inputs_from_reference(m, kind)
    lims = param_lims(inp, kind)
    th = np.concatenate([synth.draw_thetas(inp, kind, nprior, seed=seed + 1, mode='prior'),
                         synth.draw_thetas(inp, kind, nnear, seed=seed + 2, mode='near', scale=0.03)])
    th = np.concatenate([th[-1:], th[:-1]])            # row 0 = a near-truth row (base of the special rows)
    th = np.concatenate([th, special_rows(th, lims)])
    with np.errstate(all='ignore'):
        ref = np.array([fn(t.copy()) for t in th], dtype=np.float64)
    got = lf_oracle.lnprob_batch(inp, kind, th)
    assert same(ref, got), "oracle != reference for %s: max diff %g" % (
        name, np.nanmax(np.abs(np.where(np.isfinite(ref), ref - got, 0.0))))
    save = {k: v for k, v in inp.items() if v is not None}
    np.savez_compressed(os.path.join(OUT, name + '.npz'), thetas=th, lnprob_ref=ref, kind=kind, **save)
    print("%-22s kind=%-5s N=%d K=%d S=%d rows=%d finite=%d  -> bit-identical oracle" % (
        name, kind, len(inp['lum']), nfields, len(inp['zarr']), len(th), np.isfinite(ref).sum()))
    return m, inp, cat


def make_units():
    V, lf, lfz = refstub.load_reference()
    f = np.array([1e-17, 3e-17, 1e-16, 5e-16])
    L = np.array([41.0, 42.0, 42.5, 43.0, 44.0])
    out = dict(sqarcsec=V.sqarcsec, f=f, L=L,
               inv_fleming=V.inverse_fleming(3e-17, 4.56, 0.1),
               fleming_a=V.fleming(f, 3e-17, 4.56, 0.1), fleming_b=V.fleming(f, 2.72e-17, 3.5, 0.1),
               fleming_c=V.fleming(f, 3e-17, 4.56, False),
               schechter=lf.TrueLumFunc(L, -1.49, 42.5, -2.0), schechter_46=lf.TrueLumFunc(46.0, -1.49, 42.5, -2.0),
               quadcoef=np.array(lfz.getQuadCoef(42.3, 42.6, 42.7, 1.20, 1.53, 1.86)),
               schechter_z=lfz.schechter_z(L, 1.4, -1.5, 42.3, 42.6, 42.7, -2.2, -2.0, -2.1, 1.20, 1.53, 1.86))
    # SURVEY.md Appendix B lists these values; make sure the survey and this run agree
    assert out['sqarcsec'] == 42545170296.15221 and out['inv_fleming'] == 1.5301133132973276e-17
    assert same(out['fleming_a'], [0.0016099877045940335, 0.44632588920841393, 0.9610334240683099, 0.9921361596002034])
    assert same(out['quadcoef'], (-0.918273645546384, 3.41597796143255, 39.52314049586773))
    assert same(out['fleming_a'], lf_oracle.fleming(f, 3e-17, 4.56, 0.1))
    assert same(out['fleming_c'], lf_oracle.fleming(f, 3e-17, 4.56, False))
    assert same(out['schechter'], lf_oracle.schechter_log(L, -1.49, 42.5, -2.0))
    assert same(out['quadcoef'], lf_oracle.quad_coef(42.3, 42.6, 42.7, 1.20, 1.53, 1.86))
    assert same(out['schechter_z'], lf_oracle.schechter_evolving(L, 1.4, -1.5, (42.3, 42.6, 42.7), (-2.2, -2.0, -2.1),
                                                                 (1.20, 1.53, 1.86)))
    np.savez_compressed(os.path.join(OUT, 'units.npz'), **out)
    print("units                  -> bit-identical oracle")


def make_veff(name, n, nfields, seed, mcf):
    """1/V_eff weights + binned LF + bootstrap from the reference's VeffLF (lumfuncmcmc.py:515-525)."""
    V, lf, lfz = refstub.load_reference()
    import io
    import contextlib
    from scipy.integrate import quad
    cat = synth.make_catalogue(n, seed=seed, nfields=nfields)
    m = lf.LumFuncMCMC(cat['z'], flux=cat['flux'], flux_e=cat['flux_e'], Flim=list(cat['Flim']), alpha=cat['alpha'],
                       Omega_0=list(cat['Omega_0']), Flim_lims=configLF.Flim_lims, alpha_lims=configLF.alpha_lims,
                       sch_al=configLF.sch_al, Lstar=configLF.Lstar, phistar=configLF.phistar, fcmin=cat['fcmin'],
                       min_comp_frac=mcf, field_names=cat['field_names'], field_ind=cat['field_ind'],
                       nbins=20, nboot=30)
    np.random.seed(4242)
    rng_state = np.random.get_state()
    with contextlib.redirect_stdout(io.StringIO()):
        m.VeffLF()
    inp = inputs_from_reference(m, 'free')
    edges = np.linspace(min(m.lum) * 1.001, max(m.lum), m.nbins + 1)
    counts = lf_oracle.binned_lf_counts(m.lum, edges)
    vol_int = quad(m.dVdzf, m.zmin, m.zmax)[0]
    out = dict(phifunc=m.phifunc, Lavg=m.Lavg, lfbinorig=m.lfbinorig, var=m.var, edges=edges, counts=counts,
               Flims_arr=m.Flims_arr.copy(), vol_int=vol_int, sum_omega=sum(m.Omega_0), nbins=m.nbins, nboot=m.nboot,
               seed=4242)
    if mcf <= 0.001:
        phi = lf_oracle.veff_weights(m.flux, m.Flims_arr, m.alpha, m.fcmin, sum(m.Omega_0), vol_int, m.zmin)
        rel = np.max(np.abs(phi / m.phifunc - 1.0))
        assert rel < 5e-15, rel
        np.random.set_state(rng_state)
        Lavg, lfb, var, cnt, Larr = lf_oracle.boot_err_log(m.lum, m.phifunc, m.nboot, m.nbins)
        assert same(Lavg, m.Lavg) and same(lfb, m.lfbinorig) and same(var, m.var) and same(Larr, edges)
        assert np.array_equal(cnt, counts)
        print("%-22s mcf=%.2f N=%d weights rel %.1e; binned LF + bootstrap bit-identical oracle" % (name, mcf, n, rel))
    else:
        print("%-22s mcf=%.2f N=%d (per-source zmax; reference values stored)" % (name, mcf, n))
    save = {k: v for k, v in inp.items() if v is not None and k not in ('logL', 'integ_part')}
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **save, **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    make_units()
    make_case('free_k5_n2000', 'free', 2000, 5, seed=11)
    make_case('free_k3_fixal', 'free', 900, 3, seed=12, fix_sch_al=True, nprior=12, nnear=12)
    make_case('free_k2_mcf50', 'free', 600, 2, seed=13, mcf=0.5, nprior=12, nnear=12)
    make_case('fixed_k2_n800', 'fixed', 800, 2, seed=14)
    make_case('fixed_k2_fixal', 'fixed', 500, 2, seed=15, fix_sch_al=True, nprior=12, nnear=12)
    make_case('z_k2_n800', 'z', 800, 2, seed=16, evolve=(0.3, -0.2))
    make_case('z_k2_fixal', 'z', 500, 2, seed=17, fix_sch_al=True, nprior=12, nnear=12)
    make_veff('veff_k3_n400', 400, 3, seed=18, mcf=0.0)
    make_veff('veff_k2_mcf50', 250, 2, seed=19, mcf=0.5)


if __name__ == '__main__':
    main()
