"""GPU parity: the CUDA engine (through the C ABI) against the reference's golden lnprob values and the oracle.

Tolerances (BASELINE.json north_star): |gpu - ref| / |ref| <= 1e-10 for finite values in FP64; -inf <-> -inf.
"""
import numpy as np
import pytest

from lumfuncmcmc_b200 import synth
from oracle import lf_oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-10
CASES = [('free_k5_n2000', 'free'), ('free_k3_fixal', 'free'), ('free_k2_mcf50', 'free'),
         ('fixed_k2_n800', 'fixed'), ('fixed_k2_fixal', 'fixed'), ('z_k2_n800', 'z'), ('z_k2_fixal', 'z')]


def _engine(inp, kind, **kw):
    from lumfuncmcmc_b200.engine import LikelihoodEngine
    return LikelihoodEngine(inp, kind, device=0, **kw)


def _assert_parity(got, ref, rtol=RTOL):
    assert not np.isnan(got).any(), "engine returned NaN"
    assert np.array_equal(np.isneginf(got), np.isneginf(ref)), \
        "-inf sets differ at rows %s" % np.nonzero(np.isneginf(got) != np.isneginf(ref))[0]
    fin = np.isfinite(ref)
    rel = np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])
    assert rel.max() <= rtol, "max rel diff %.3e at row %d" % (rel.max(), np.nonzero(fin)[0][rel.argmax()])
    return rel.max()


@pytest.mark.parametrize('name,kind', CASES)
@pytest.mark.parametrize('literal', [False, True])
def test_golden_lnprob(golden, name, kind, literal):
    g = golden(name)
    eng = _engine(g, kind, force_literal=literal)
    got = eng.lnprob(g['thetas'])
    info = eng.last_call_info()
    _assert_parity(got, g['lnprob_ref'])
    assert info['rejected'] + info['fast'] + info['literal'] == len(g['thetas'])
    if literal:
        assert info['fast'] == 0
    else:
        assert info['fast'] > 0          # the optimised kernels are the ones exercised
    eng.close()


@pytest.mark.parametrize('name,kind', [('free_k5_n2000', 'free'), ('z_k2_n800', 'z'), ('fixed_k2_n800', 'fixed')])
def test_scalar_call_equals_batch_row(golden, name, kind):
    g = golden(name)
    eng = _engine(g, kind)
    batch = eng.lnprob(g['thetas'][:6])
    for i in range(6):
        one = eng.lnprob(g['thetas'][i])
        assert isinstance(one, float)
        assert one == batch[i] or (np.isneginf(one) and np.isneginf(batch[i]))
    eng.close()


def test_device_api_matches_host_api(golden):
    import torch
    g = golden('free_k5_n2000')
    eng = _engine(g, 'free')
    host = eng.lnprob(g['thetas'])
    d_th = torch.from_numpy(np.ascontiguousarray(g['thetas'])).cuda()
    d_out = eng.lnprob_device(d_th)
    torch.cuda.synchronize()
    dev = d_out.cpu().numpy()
    assert np.array_equal(host, dev, equal_nan=True)
    eng.close()


def _shard(inp, lo_frac, hi_frac):
    """Take the same fraction of every field (source sharding keeps field membership)."""
    fi = np.asarray(inp['field_ind'])
    keep, new_fi = [], [0]
    for k in range(len(fi) - 1):
        n = fi[k + 1] - fi[k]
        a, b = fi[k] + int(n * lo_frac), fi[k] + int(n * hi_frac)
        keep.append(np.arange(a, b))
        new_fi.append(new_fi[-1] + (b - a))
    keep = np.concatenate(keep)
    out = dict(inp)
    for key in ('lum', 'z', 'Om_arr', 'flux'):
        if key in inp and inp[key] is not None:
            out[key] = np.asarray(inp[key])[keep]
    out['field_ind'] = np.array(new_fi, dtype=np.int64)
    return out


@pytest.mark.parametrize('name,kind', [('free_k5_n2000', 'free'), ('z_k2_n800', 'z'), ('fixed_k2_n800', 'fixed')])
def test_additivity_over_source_shards(golden, name, kind):
    """lnprob is a plain sum over sources: two shards (each integrating half of the walkers) add up to the whole."""
    g = golden(name)
    th = g['thetas']
    full = _engine(g, kind)
    whole = full.lnprob(th)
    parts = []
    for r, (a, b) in enumerate([(0.0, 0.37), (0.37, 1.0)]):
        e = _engine(_shard(g, a, b), kind, quadrature_share=(r, 2))
        parts.append(e.lnprob(th))
        e.close()
    with np.errstate(invalid='ignore'):
        total = parts[0] + parts[1]
    _assert_parity(total, whole, rtol=1e-12)
    _assert_parity(total, g['lnprob_ref'])
    full.close()


def test_permutation_invariance_within_field(golden):
    g = golden('free_k5_n2000')
    rng = np.random.default_rng(3)
    fi = g['field_ind']
    perm = np.concatenate([fi[k] + rng.permutation(fi[k + 1] - fi[k]) for k in range(len(fi) - 1)])
    p = dict(g)
    p['lum'], p['z'] = g['lum'][perm], g['z'][perm]
    a, b = _engine(g, 'free'), _engine(p, 'free')
    _assert_parity(b.lnprob(g['thetas']), a.lnprob(g['thetas']), rtol=1e-13)
    a.close()
    b.close()


@pytest.mark.parametrize('kind', ['free', 'fixed', 'z'])
def test_midsize_synthetic_against_oracle(kind):
    """N = 2e5 sources, W = 96 walkers (mixed near-truth / prior draws); oracle on every walker."""
    cat = synth.make_catalogue(200000, seed=21, evolve=(0.3, -0.2) if kind == 'z' else None)
    inp = synth.direct_inputs(cat, nknots=2048, size_ln=101 if kind == 'free' else 201, tabulated=(kind != 'free'))
    th = np.concatenate([synth.draw_thetas(inp, kind, 64, seed=5, mode='near', scale=0.02),
                         synth.draw_thetas(inp, kind, 32, seed=6, mode='prior')])
    eng = _engine(inp, kind)
    got = eng.lnprob(th)
    ref = lf_oracle.lnprob_batch(inp, kind, th)
    rel = _assert_parity(got, ref)
    info = eng.last_call_info()
    assert info['fast'] >= 64
    print(kind, 'max rel', rel, info)
    eng.close()


def test_plain_fleming_curve_when_fcmin_is_zero():
    cat = synth.make_catalogue(5000, seed=8, nfields=3, fcmin=0.0)
    inp = synth.direct_inputs(cat, nknots=512)
    assert inp['fcmin'] == 0.0
    th = np.concatenate([synth.draw_thetas(inp, 'free', 24, seed=1), synth.draw_thetas(inp, 'free', 8, seed=2, mode='prior')])
    for literal in (False, True):
        eng = _engine(inp, 'free', force_literal=literal)
        _assert_parity(eng.lnprob(th), lf_oracle.lnprob_batch(inp, 'free', th))
        eng.close()


def test_empty_and_ragged_fields():
    """A field with no sources and a single-source field (ragged field_ind) are legal inputs."""
    cat = synth.make_catalogue(400, seed=9, nfields=3)
    inp = synth.direct_inputs(cat, nknots=256)
    fi = inp['field_ind']
    keep = np.concatenate([np.arange(fi[0], fi[1]), np.arange(fi[2], fi[2] + 1)])      # field 1 empty, field 2 one source
    inp = dict(inp)
    inp['lum'], inp['z'] = inp['lum'][keep], inp['z'][keep]
    inp['field_ind'] = np.array([0, fi[1], fi[1], fi[1] + 1], dtype=np.int64)
    th = synth.draw_thetas(inp, 'free', 16, seed=1)
    eng = _engine(inp, 'free')
    _assert_parity(eng.lnprob(th), lf_oracle.lnprob_batch(inp, 'free', th))
    eng.close()


def test_veff_weights_and_bit_exact_bin_counts(golden):
    g = golden('veff_k3_n400')
    from lumfuncmcmc_b200.engine import LikelihoodEngine
    eng = LikelihoodEngine(golden('free_k3_fixal'), 'free', device=0)
    K = len(g['Flim'])
    phi, counts, sums = eng.veff_bin(g['flux'], g['lum'], g['field_ind'], g['Flim'], g['alpha'], g['fcmin'],
                                     g['sum_omega'], g['vol_int'], g['edges'])
    np.testing.assert_allclose(phi, g['phifunc'], rtol=1e-13)
    assert np.array_equal(counts, g['counts'])                      # integers: bit-exact
    dL = g['Lavg'][1] - g['Lavg'][0]
    np.testing.assert_allclose(sums / dL, g['lfbinorig'], rtol=1e-12)
    # one bootstrap replicate with the reference's RNG stream (VmaxLumFunc.py:353)
    np.random.seed(int(g['seed']))
    boot = np.random.randint(len(phi), size=len(phi))
    mult = np.bincount(boot, minlength=len(phi))
    bc, bs = eng.boot_bin(mult)
    Lb, pb = g['lum'][boot], g['phifunc'][boot]
    want_c = lf_oracle.binned_lf_counts(Lb, g['edges'])
    assert np.array_equal(bc, want_c)
    want_s = np.array([pb[(Lb >= g['edges'][j]) & (Lb < g['edges'][j + 1])].sum() for j in range(len(want_c))])
    np.testing.assert_allclose(bs, want_s, rtol=1e-12)
    assert K == 3
    eng.close()


@pytest.mark.parametrize('nbins', [50, 700])
def test_veff_bit_exact_counts_large(nbins):
    """1e6 sources: integer counts identical to NumPy's mask counts on the exact linspace edges.  50 bins (the
    reference's default) runs the lane-private histogram kernel, 700 bins the shared-memory-atomics fallback."""
    rng = np.random.default_rng(4)
    n = 1000000
    lum = rng.uniform(40.9, 44.0, n)
    lum[:1000] = lum[1000:2000]                    # ties
    flux = 10 ** rng.uniform(-17.2, -14.5, n)
    fi = np.array([0, n // 3, n // 2, n], dtype=np.int64)
    edges = np.linspace(lum.min() * 1.001, lum.max(), nbins + 1)
    lum[5000:5050] = edges[:50]                    # values exactly on the edges
    from lumfuncmcmc_b200.engine import LikelihoodEngine
    eng = LikelihoodEngine(synth.direct_inputs(synth.make_catalogue(300, seed=1, nfields=3), nknots=64), 'free', device=0)
    flim = [2.72, 3.61, 2.55]
    phi, counts, sums = eng.veff_bin(flux, lum, fi, flim, 4.56, 0.1, 1.0e6, 3.0e10, edges)
    want = np.histogram(lum[(lum >= edges[0]) & (lum < edges[-1])], bins=edges)[0]
    # np.histogram closes the last bin on the right; with the mask above both conventions agree
    assert np.array_equal(counts, want)
    flims_arr = np.repeat(flim, np.diff(fi))
    ref_phi = lf_oracle.veff_weights(flux, flims_arr, 4.56, 0.1, 1.0e6, 3.0e10, 0.0)
    np.testing.assert_allclose(phi, ref_phi, rtol=1e-13)
    j = np.searchsorted(edges, lum, side='right') - 1
    ok = (lum >= edges[0]) & (lum < edges[-1])
    want_s = np.bincount(j[ok], weights=ref_phi[ok], minlength=nbins)[:nbins]
    np.testing.assert_allclose(sums, want_s, rtol=1e-12)
    # a bootstrap replicate on the resident sample: counts exact, sums to 1e-12
    mult = np.bincount(rng.integers(0, n, n), minlength=n)
    bc, bs = eng.boot_bin(mult)
    assert np.array_equal(bc, np.bincount(j[ok], weights=mult[ok], minlength=nbins)[:nbins].astype(np.int64))
    np.testing.assert_allclose(bs, np.bincount(j[ok], weights=(ref_phi * mult)[ok], minlength=nbins)[:nbins], rtol=1e-12)
    eng.close()


@pytest.mark.parametrize('fcmin', [0.1, 0.0])
@pytest.mark.parametrize('edge_kind', ['linspace', 'irregular'])
def test_veff_routes_agree_with_numpy(fcmin, edge_kind):
    """The streaming kernel's lean route (complete single-field trips, near-uniform edges) and its general route
    (irregular edges, field boundaries, tails) against NumPy: counts bit-exact, weights 1e-13; plain Fleming
    (fcmin = 0), sources outside the fast-math range (literal fall-back), non-finite and out-of-range luminosities."""
    rng = np.random.default_rng(11)
    n = 600000 + 77
    nbins = 50
    lum = rng.uniform(40.5, 44.5, n)
    flux = 10 ** rng.uniform(-17.5, -14.0, n)
    if edge_kind == 'linspace':
        edges = np.linspace(41.0, 44.0, nbins + 1)
    else:
        edges = np.sort(np.concatenate([[41.0, 44.0], rng.uniform(41.0, 44.0, nbins - 1)]))
    # out-of-range luminosities carry the extreme fluxes, so the per-bin sums stay finite
    lum[:64] = 40.0
    flux[:64] = 10 ** rng.uniform(-24.0, -19.0, 64)            # decay argument < 1e-6 / |ln comp| > 690: literal route
    flux[64:96] = 10 ** rng.uniform(-13.0, -11.0, 32)          # f / f_tau > 690
    lum[64:96] = rng.uniform(41.0, 44.0, 32)
    lum[100:104] = [np.nan, np.inf, -np.inf, 1e300]
    lum[200:200 + nbins + 1] = edges                           # exactly on every edge
    lum[300:300 + nbins + 1] = np.nextafter(edges, -np.inf)    # one ulp below every edge
    fi = np.array([0, 1000, 1000, 250001, n], dtype=np.int64)  # an empty field, boundaries inside trips
    flim = [2.72, 3.61, 2.55, 3.31]
    from lumfuncmcmc_b200.engine import LikelihoodEngine
    eng = LikelihoodEngine(synth.direct_inputs(synth.make_catalogue(300, seed=1, nfields=3), nknots=64), 'free', device=0)
    phi, counts, sums = eng.veff_bin(flux, lum, fi, flim, 4.56, fcmin, 1.0e6, 3.0e10, edges)
    with np.errstate(invalid='ignore'):
        ok = (lum >= edges[0]) & (lum < edges[-1])
    j = np.searchsorted(edges, lum[ok], side='right') - 1
    want = np.bincount(j, minlength=nbins)[:nbins]
    assert np.array_equal(counts, want)
    assert counts.sum() == ok.sum()
    flims_arr = np.repeat(flim, np.diff(fi))
    with np.errstate(all='ignore'):
        ref_phi = lf_oracle.veff_weights(flux, flims_arr, 4.56, fcmin, 1.0e6, 3.0e10, 0.0)
    fin = np.isfinite(ref_phi)
    assert np.array_equal(np.isfinite(phi), fin)
    # fc = (1 + n / sqrt(1 + n^2)) / 2 cancels for faint sources: NumPy's own value carries eps / fc (times the
    # exponent 1 / (1 - e^-x) of the modified form), and so does any other evaluation order
    with np.errstate(all='ignore'):
        nn = 4.56 * np.log10(flux / (1.0e-17 * flims_arr))
        cond = 1.0 / (0.5 * (1.0 + nn / np.sqrt(1.0 + nn * nn)))
        if fcmin:
            cond = cond / -np.expm1(-flux / (1.0e-17 * flims_arr) * 10 ** (np.sqrt(abs((2 * fcmin - 1) ** 2 / (1 - (2 * fcmin - 1) ** 2))) / 4.56))
    assert np.all(np.abs(phi[fin] - ref_phi[fin]) <= (1e-13 + 1e-15 * cond[fin]) * ref_phi[fin])
    typical = fin & (cond < 50.0)
    assert typical.sum() > 0.8 * n
    np.testing.assert_allclose(phi[typical], ref_phi[typical], rtol=1e-13)
    np.testing.assert_allclose(sums, np.bincount(j, weights=ref_phi[ok], minlength=nbins)[:nbins], rtol=1e-12)
    # the resident rows feed the bootstrap replicates
    mult = np.bincount(rng.integers(0, n, n), minlength=n)
    bc, bs = eng.boot_bin(mult)
    assert np.array_equal(bc, np.bincount(j, weights=mult[ok], minlength=nbins)[:nbins].astype(np.int64))
    np.testing.assert_allclose(bs, np.bincount(j, weights=(ref_phi * mult)[ok], minlength=nbins)[:nbins], rtol=1e-12)
    # binning caller-provided weights (getBootErrLog's entry point) takes the same routes
    c2, s2 = eng.bin_weights(lum, np.where(fin, ref_phi, 0.0), edges)
    assert np.array_equal(c2, want)
    np.testing.assert_allclose(s2, np.bincount(j, weights=np.where(fin, ref_phi, 0.0)[ok], minlength=nbins)[:nbins], rtol=1e-12)
    eng.close()


# ---------------------------------------------------------------------------------------------------------------
# optional FP32 mode of the walker x source loop: 1e-5 relative (BASELINE.json north_star)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name,kind', [('free_k5_n2000', 'free'), ('free_k3_fixal', 'free'), ('z_k2_n800', 'z')])
def test_fp32_mode_golden(golden, name, kind):
    g = golden(name)
    eng = _engine(g, kind, precision='f32')
    got = eng.lnprob(g['thetas'])
    rel = _assert_parity(got, g['lnprob_ref'], rtol=1e-5)
    assert eng.last_call_info()['fast'] > 0
    # and it really is a different arithmetic: not bit-identical to the FP64 engine
    e64 = _engine(g, kind)
    ref64 = e64.lnprob(g['thetas'])
    fin = np.isfinite(ref64)
    assert np.any(got[fin] != ref64[fin])
    print(name, 'fp32 max rel', rel)
    eng.close()
    e64.close()


@pytest.mark.parametrize('kind', ['free', 'z'])
def test_fp32_mode_midsize_against_oracle(kind):
    cat = synth.make_catalogue(200000, seed=22, evolve=(0.3, -0.2) if kind == 'z' else None)
    inp = synth.direct_inputs(cat, nknots=2048, size_ln=101 if kind == 'free' else 201, tabulated=(kind != 'free'))
    th = np.concatenate([synth.draw_thetas(inp, kind, 48, seed=5, mode='near', scale=0.02),
                         synth.draw_thetas(inp, kind, 16, seed=6, mode='prior')])
    eng = _engine(inp, kind, precision='f32')
    got = eng.lnprob(th)
    rel = _assert_parity(got, lf_oracle.lnprob_batch(inp, kind, th), rtol=1e-5)
    print(kind, 'fp32 max rel', rel, eng.last_call_info())
    eng.close()


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json full size (1e7 sources): oracle on a few walkers + size-independent properties on the ensemble
# ---------------------------------------------------------------------------------------------------------------
def test_full_size_catalogue_properties_and_oracle_spot_check():
    n = 10_000_000
    cat = synth.make_catalogue(n, seed=77)
    inp = synth.direct_inputs(cat, nknots=4096, size_ln=101)
    th = np.concatenate([synth.draw_thetas(inp, 'free', 60, seed=5, mode='near', scale=0.02),
                         synth.draw_thetas(inp, 'free', 4, seed=6, mode='prior')])
    eng = _engine(inp, 'free')
    whole = eng.lnprob(th)
    assert eng.last_call_info()['fast'] >= 60
    # (1) the reference algorithm on ten walkers, two of them prior draws (1e8 terms of NumPy work)
    sel = [0, 7, 15, 23, 31, 39, 47, 59, 60, 63]
    ref = lf_oracle.lnprob_batch(inp, 'free', th[sel])
    _assert_parity(whole[sel], ref)
    eng.close()
    # (1b) the optional FP32 loop at the same size: 1e-5 against the FP64 engine on every walker, and against the oracle
    e32 = _engine(inp, 'free', precision='f32')
    got32 = e32.lnprob(th)
    _assert_parity(got32, whole, rtol=1e-5)
    _assert_parity(got32[sel], ref, rtol=1e-5)
    e32.close()
    # (2) additivity over three unequal source shards, each integrating a third of the walkers
    parts = []
    for r, (a, b) in enumerate([(0.0, 0.21), (0.21, 0.64), (0.64, 1.0)]):
        e = _engine(_shard(inp, a, b), 'free', quadrature_share=(r, 3))
        parts.append(e.lnprob(th))
        e.close()
    with np.errstate(invalid='ignore'):
        total = parts[0] + parts[1] + parts[2]
    _assert_parity(total, whole, rtol=1e-12)
    # (3) permutation of the sources inside every field leaves the result unchanged (sums reorder: 1e-13)
    rng = np.random.default_rng(9)
    fi = inp['field_ind']
    perm = np.concatenate([fi[k] + rng.permutation(fi[k + 1] - fi[k]) for k in range(len(fi) - 1)])
    p = dict(inp)
    p['lum'], p['z'] = inp['lum'][perm], inp['z'][perm]
    e = _engine(p, 'free')
    _assert_parity(e.lnprob(th), whole, rtol=1e-13)
    e.close()


def test_config3_size_z_model_against_oracle():
    """BASELINE.json configs[2] at its own size: redshift-evolving model, 1e6 sources x 512 walkers; the oracle on twelve of
    them (prior draws included), the -inf sets of the whole ensemble against the literal kernels."""
    n = 1_000_000
    cat = synth.make_catalogue(n, seed=78, evolve=(0.3, -0.2))
    inp = synth.direct_inputs(cat, nknots=4096, size_ln=201, tabulated=True)
    th = np.concatenate([synth.draw_thetas(inp, 'z', 480, seed=5, mode='near', scale=0.02),
                         synth.draw_thetas(inp, 'z', 32, seed=6, mode='prior')])
    eng = _engine(inp, 'z')
    got = eng.lnprob(th)
    assert eng.last_call_info()['fast'] >= 480
    sel = [0, 100, 200, 300, 400, 479, 480, 485, 490, 495, 505, 511]
    ref = lf_oracle.lnprob_batch(inp, 'z', th[sel])
    _assert_parity(got[sel], ref)
    eng.close()
    lit = _engine(inp, 'z', force_literal=True)
    sub = np.r_[0:16, 480:512]
    _assert_parity(lit.lnprob(th[sub]), got[sub], rtol=1e-10)
    lit.close()
    e32 = _engine(inp, 'z', precision='f32')
    _assert_parity(e32.lnprob(th), got, rtol=1e-5)
    e32.close()


@pytest.mark.gpu
def test_z_model_walker_with_a_far_away_middle_pivot_takes_the_literal_kernels():
    """The z-evolving source loop evaluates 2^(log2(10) (lum - L*(z)) - log2(10) (42 - L*(z2))) without clamping the exponent
    field; k_prologue keeps walkers whose middle-pivot value L*(z2) lies more than 200 dex from 42 out of the fast class.  With
    the middle pivot far outside the catalogue's redshift range L*(z) is perfectly ordinary over the sources while L*(z2) is
    not: such a walker must come out right (literal kernels), and its neighbours must stay fast."""
    cat = synth.make_catalogue(20000, seed=91, evolve=(0.3, -0.2))
    inp = dict(synth.direct_inputs(cat, nknots=1024, size_ln=101, tabulated=True))
    inp['z2'] = 60.0                                     # pivots: z1, z3 at the ends of the survey, z2 far away
    inp['Lstar_lims'] = [-500.0, 900.0]
    th = synth.draw_thetas(inp, 'z', 40, seed=3, mode='near', scale=0.01)
    z1, z3 = float(inp['z1']), float(inp['z3'])
    # L*(z) linear from 42.4 to 42.6 across the survey, extrapolated to z2 and then bent so that L*(z2) is 450 / 800
    for row, L2 in ((5, 450.0), (17, 800.0)):
        th[row, 0], th[row, 2], th[row, 1] = 42.4, 42.6, L2
    for row in range(40):
        if row not in (5, 17):                           # the others: the straight line itself (an ordinary L*(z2))
            th[row, 1] = th[row, 0] + (th[row, 2] - th[row, 0]) * (60.0 - z1) / (z3 - z1)
    ref = lf_oracle.lnprob_batch(inp, 'z', th)
    assert np.isfinite(ref[[5, 17]]).all()
    eng = _engine(inp, 'z')
    got = eng.lnprob(th)
    info = eng.last_call_info()
    _assert_parity(got, ref)
    assert info['literal'] >= 2 and info['fast'] >= 30, info
    eng.close()


# ---------------------------------------------------------------------------------------------------------------
# compressed catalogue (opt-in): weighted pseudo-sources instead of the walker x source loop
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['free_k5_n2000', 'free_k3_fixal', 'free_k2_mcf50'])
def test_compressed_catalogue_golden(golden, name):
    g = golden(name)
    eng = _engine(g, 'free', compress=True)
    assert 0 < eng.npseudo
    got = eng.lnprob(g['thetas'])
    _assert_parity(got, g['lnprob_ref'])
    eng.uncompress_catalogue()
    _assert_parity(got, eng.lnprob(g['thetas']), rtol=1e-12)
    eng.close()


def test_compressed_catalogue_midsize_against_brute_force_and_oracle():
    """N = 2e5 -> a few thousand pseudo-sources; walkers near the truth and over the whole prior box (alpha_c up to 7)."""
    cat = synth.make_catalogue(200000, seed=23)
    inp = synth.direct_inputs(cat, nknots=2048, size_ln=101)
    th = np.concatenate([synth.draw_thetas(inp, 'free', 64, seed=5, mode='near', scale=0.02),
                         synth.draw_thetas(inp, 'free', 64, seed=6, mode='prior')])
    th[64:72, -1] = 7.0                                          # the prior's steepest completeness curve
    brute, comp = _engine(inp, 'free'), _engine(inp, 'free', compress=True)
    assert comp.npseudo < 20000
    a, b = brute.lnprob(th), comp.lnprob(th)
    rel = _assert_parity(b, a, rtol=1e-12)
    _assert_parity(b[:16], lf_oracle.lnprob_batch(inp, 'free', th[:16]))
    print('compressed: %d pseudo-sources for %d sources, max rel diff vs brute force %.2e' % (comp.npseudo, 200000, rel))
    # without the prior gate a walker may sit above the alpha_c the weights were built for: it must take the literal
    # kernels on the real sources, not the interpolant
    steep = th[:4].copy()
    steep[:, -1] = 9.5
    _assert_parity(comp.lnlike(steep), brute.lnlike(steep), rtol=1e-12)
    assert comp.last_call_info()['literal'] == 4
    brute.close()
    comp.close()


def test_peer_memory_allreduce_degenerates_to_identity_on_one_rank(golden):
    """world = 1: the peer-memory exchange kernel (stores, flag, wait, rank-ordered sum) returns the vector unchanged,
    for both buffer parities and for lengths that are not a multiple of the 256-walker chunk.  The multi-rank case is
    exercised by tests/run_multi_gpu_parity.py under torchrun (profiles/r01_multi_gpu_parity_2gpu_p2p.log)."""
    import torch
    g = golden('free_k3_fixal')
    eng = _engine(g, 'free')
    handle = eng.peer_buffer_create(0, 1, 1000)
    assert len(handle) == 64
    eng.peer_buffer_connect([handle])
    for n in (1, 255, 256, 257, 1000):
        v = torch.arange(n, dtype=torch.float64, device='cuda') * 0.5 - 3.0
        v[0] = float('-inf')
        want = v.clone()
        for _ in range(3):
            eng.allreduce_device(v)
        torch.cuda.synchronize()
        assert torch.equal(v, want)
    assert not eng.peer_timed_out()
    with pytest.raises(Exception):
        eng.allreduce_device(torch.zeros(5000, dtype=torch.float64, device='cuda'))
    eng.close()


def test_peer_exchange_timeout_poisons_the_result_and_is_sticky(golden):
    """A rank that never arrives: the wait expires (bound set to 20 ms here), the exchange returns NaN -- never a stale
    partial sum --, the flag stays set, later exchanges are refused until lf_peer_reset.  World = 2 with no peer mapped
    for rank 1 (all-zero handle row): its flag is never raised."""
    import torch
    g = golden('free_k3_fixal')
    eng = _engine(g, 'free')
    handle = eng.peer_buffer_create(0, 2, 512)
    eng.peer_buffer_connect([handle, bytes(64)])
    eng.peer_set_timeout(0.02)
    v = torch.arange(300, dtype=torch.float64, device='cuda') + 1.0
    eng.allreduce_device(v)
    torch.cuda.synchronize()
    assert torch.isnan(v).all()
    assert eng.peer_timed_out() and eng.peer_timed_out()          # sticky
    with pytest.raises(Exception, match='timed out'):
        eng.allreduce_device(v)
    eng.peer_reset()
    assert not eng.peer_timed_out()
    eng.close()


def test_lnprob_on_a_caller_stream_is_ordered_against_the_context_stream(golden):
    """The per-call scratch is shared by every entry point of a context: a device call on the caller's stream followed
    at once by host-API calls (context stream) and the reverse must give the same numbers as serial execution."""
    import torch
    g = golden('free_k5_n2000')
    eng = _engine(g, 'free')
    th = np.ascontiguousarray(g['thetas'])
    want = eng.lnprob(th)
    d_th = torch.from_numpy(th).cuda()
    side = torch.cuda.Stream()
    outs = []
    for rep in range(20):
        with torch.cuda.stream(side):
            outs.append(eng.lnprob_device(d_th, stream=side))
        got_host = eng.lnprob(th[::-1].copy())                    # context stream, right behind the side-stream call
        assert np.array_equal(got_host, want[::-1], equal_nan=True)
    torch.cuda.synchronize()
    for o in outs:
        assert np.array_equal(o.cpu().numpy(), want, equal_nan=True)
    eng.close()


def test_device_resampled_bootstrap_equals_host_replay_of_the_philox_stream():
    """rng='device': every replicate's multiplicities come from a Philox stream on the GPU; a host replay of the same
    stream gives the same integer counts, sums agree to 1e-12, and the variances are statistically those of the
    reference's host-RNG bootstrap."""
    from lumfuncmcmc_b200 import VmaxLumFunc as V
    from lumfuncmcmc_b200.engine import VeffEngine
    from lumfuncmcmc_b200.sampler import philox4x32_10
    rng = np.random.default_rng(12)
    n, nb, seed = 200003, 30, 0xfeedbeef12345678
    lum = rng.uniform(41.0, 43.5, n)
    phi = 10 ** rng.uniform(-7, -5, n)
    edges = np.linspace(lum.min() * 1.001, lum.max(), nb + 1)
    eng = VeffEngine()
    counts0, sums0 = eng.bin_weights(lum, phi, edges)
    j = np.searchsorted(edges, lum, side='right') - 1
    ok = (lum >= edges[0]) & (lum < edges[-1])
    for rep in (0, 7):
        bc, bs = eng.boot_bin_device(seed, rep)
        q = np.arange((n + 3) // 4, dtype=np.uint64)
        r = philox4x32_10(q & np.uint64(0xFFFFFFFF), q >> np.uint64(32), np.full(len(q), rep), np.zeros(len(q)), seed & 0xFFFFFFFF, seed >> 32)
        draws = np.stack(r, axis=1).ravel()[:n].astype(np.uint64)
        mult = np.bincount(((draws * np.uint64(n)) >> np.uint64(32)).astype(np.int64), minlength=n)
        assert mult.sum() == n
        assert np.array_equal(bc, np.bincount(j[ok], weights=mult[ok], minlength=nb)[:nb].astype(np.int64))
        np.testing.assert_allclose(bs, np.bincount(j[ok], weights=(phi * mult)[ok], minlength=nb)[:nb], rtol=1e-12)
        bc2, _ = eng.boot_bin_device(seed, rep)
        assert np.array_equal(bc, bc2)                                   # reproducible
    np.random.seed(3)
    _, lf_h, var_h = V.getBootErrLog(lum, phi, 1.2, 1.9, nboot=200, nbin=nb, Larr=edges, engine=eng)
    _, lf_d, var_d = V.getBootErrLog(lum, phi, 1.2, 1.9, nboot=200, nbin=nb, Larr=edges, engine=eng, rng='device', seed=seed)
    assert np.array_equal(lf_h, lf_d)
    # two independent 200-replicate variance estimates: each bin within a factor 2, the average over bins within 10 %
    assert np.all(np.abs(np.log(var_d / var_h)) < np.log(2.0)) and abs(np.mean(var_d / var_h) - 1.0) < 0.1
    eng.close()


def test_compressed_catalogue_z_model(golden):
    g = golden('z_k2_n800')
    eng = _engine(g, 'z', compress=True)
    assert eng.npseudo > 0
    _assert_parity(eng.lnprob(g['thetas']), g['lnprob_ref'])
    eng.close()
    cat = synth.make_catalogue(200000, seed=24, evolve=(0.3, -0.2))
    inp = synth.direct_inputs(cat, nknots=2048, size_ln=201, tabulated=True)
    th = np.concatenate([synth.draw_thetas(inp, 'z', 64, seed=5, mode='near', scale=0.02),
                         synth.draw_thetas(inp, 'z', 64, seed=6, mode='prior')])
    brute, comp = _engine(inp, 'z'), _engine(inp, 'z', compress=True)
    a, b = brute.lnprob(th), comp.lnprob(th)
    rel = _assert_parity(b, a, rtol=1e-12)
    info = comp.last_call_info()
    assert info['fast'] >= 64
    _assert_parity(b[:12], lf_oracle.lnprob_batch(inp, 'z', th[:12]))
    print('z compressed: %d pseudo-sources, max rel diff vs brute force %.2e, classes %s' % (comp.npseudo, rel, info))
    brute.close()
    comp.close()
