"""Synthetic flux-limited emission-line catalogues drawn from known Schechter parameters.

Catalogue recipe: K fields with the reference's configured areas and 50%-completeness
fluxes (reference configLF.py:6,10,18-21,31-32), z ~ dV/dz on [zmin, zmax], logL from a Schechter function
with (logL*, logphi*, alpha) thinned by the modified Fleming completeness in flux -- i.e. sources are drawn
directly from the *observed* density  Phi(logL) * fleming(flux(logL, z)) * dV/dz  by inverse-CDF on a fine
(z, logL) grid.  Everything is seeded and host-side (NumPy); it feeds tests and bench only.

``direct_inputs`` additionally builds the arrays the likelihood engine consumes ("engine inputs") without
going through the class constructor, for catalogue sizes where the reference's set-up chain would take hours
(SURVEY.md section 8c "Large-N oracle").  It follows the reference's set-up arithmetic
(lumfuncmcmc.py:180-235) with a small-knot D_L interpolant and direct (not spline-tabulated) Omega.
"""
import numpy as np

from . import configLF
from .cosmology import cosmo

SQARCSEC = (180. / np.pi * 3600.0) ** 2

#: CUDA ordinal used for the O(N) cosmology passes of large catalogues when a GPU is present (bench / tests set it to the
#: rank's device); below ``setup_gpu.GPU_MIN_POINTS`` points, or without a GPU, the host arithmetic of cosmology.py runs
DEVICE = 0


def _lumdist(z):
    """D_L [Mpc]: ``cosmology.cosmo`` on the host, or the same arithmetic on the GPU (``lf_cosmo_distances``, equal to
    ~2e-15) for large arrays -- synthetic data either way, and both the oracle and the engine consume what comes out."""
    from .setup_gpu import GPU_MIN_POINTS, cosmo_distances, gpu_count
    z = np.asarray(z, dtype=np.float64)
    if z.size >= GPU_MIN_POINTS and gpu_count() > 0:
        return cosmo_distances(cosmo, z, device=DEVICE, want_dv=False)[0]
    return cosmo.luminosity_distance(z)


def _fleming(f, f50, alpha, fcmin):
    n = alpha * np.log10(f / f50)
    fc = 0.5 * (1.0 + n / np.sqrt(1.0 + n * n))
    if not fcmin:
        return fc
    a = (2.0 * fcmin - 1.0) ** 2
    ftau = f50 * 10.0 ** (-np.sqrt(abs(a / (1.0 - a))) / alpha)
    return fc ** (1.0 / (1.0 - np.exp(-f / ftau)))


def make_catalogue(n_sources, seed=0, nfields=5, Lstar=42.5, phistar=-2.0, sch_al=-1.49, Flim=None, alpha=None,
                   Omega_0=None, fcmin=None, zmin=1.16, zmax=1.90, Lc=40.0, Lh=46.0, evolve=None,
                   flux_err_frac=0.1):
    """Draw a catalogue.  Returns dict with per-field lists ``z``, ``flux``, ``flux_e`` (flux in 1e-17 cgs, the
    unit the reference's constructor expects), ``field_ind`` and the generating parameters.

    ``evolve=(dLstar_dz, dphistar_dz)`` makes logL*, logphi* linear in (z - zmid) for the z-evolving model.
    """
    rng = np.random.default_rng(seed)
    Flim = list(configLF.Flim if Flim is None else Flim)[:nfields]
    Omega_0 = list(configLF.Omega_0 if Omega_0 is None else Omega_0)[:nfields]
    alpha = configLF.alpha if alpha is None else alpha
    fcmin = configLF.fcmin if fcmin is None else fcmin
    nz, nl = 96, 1536
    zc = zmin + (zmax - zmin) * (np.arange(nz) + 0.5) / nz
    lc = Lc + (Lh - Lc) * (np.arange(nl) + 0.5) / nl
    DL = cosmo.luminosity_distance(zc)
    dV = cosmo.differential_comoving_volume(zc)
    fluxgrid = 10.0 ** lc[None, :] / (4.0 * np.pi * (3.086e24 * DL[:, None]) ** 2)
    zmid = 0.5 * (zmin + zmax)
    if evolve is None:
        Ls, ps = np.full(nz, Lstar), np.full(nz, phistar)
    else:
        Ls, ps = Lstar + evolve[0] * (zc - zmid), phistar + evolve[1] * (zc - zmid)
    dex = lc[None, :] - Ls[:, None]
    phi = np.log(10.0) * 10.0 ** ps[:, None] * 10.0 ** (dex * (sch_al + 1.0)) * np.exp(-10.0 ** dex)
    dens, expected = [], []
    for k in range(nfields):
        d = phi * dV[:, None] * (Omega_0[k] / SQARCSEC) * _fleming(fluxgrid, 1.0e-17 * Flim[k], alpha, fcmin)
        dens.append(d)
        expected.append(d.sum() * (zmax - zmin) / nz * (Lh - Lc) / nl)
    expected = np.array(expected)
    counts = np.floor(n_sources * expected / expected.sum()).astype(np.int64)
    counts[np.argmax(counts)] += n_sources - counts.sum()
    z_l, f_l, fe_l, dl_l = [], [], [], []
    for k in range(nfields):
        cdf = np.cumsum(dens[k].ravel())
        cdf /= cdf[-1]
        cell = np.searchsorted(cdf, rng.random(counts[k]), side='right')
        cell = np.minimum(cell, nz * nl - 1)
        iz, il = np.divmod(cell, nl)
        z = zmin + (zmax - zmin) * (iz + rng.random(counts[k])) / nz
        logL = Lc + (Lh - Lc) * (il + rng.random(counts[k])) / nl
        dl = _lumdist(z)
        flux = 10.0 ** logL / (4.0 * np.pi * (3.086e24 * dl) ** 2) / 1.0e-17
        dl_l.append(dl)
        z_l.append(z)
        f_l.append(flux)
        fe_l.append(flux_err_frac * flux)
    field_ind = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return dict(z=z_l, flux=f_l, flux_e=fe_l, field_ind=field_ind,
                field_names=np.array(['F%d' % k for k in range(nfields)]),
                Flim=Flim, alpha=alpha, Omega_0=Omega_0, fcmin=fcmin, expected_total=float(expected.sum()), _DL=dl_l,
                truth=dict(Lstar=Lstar, phistar=phistar, sch_al=sch_al, evolve=evolve))


def direct_inputs(cat, nknots=4096, size_ln=101, Lh=46.0, z_pivots=(1.20, 1.53, 1.86), tabulated=False, zrange=None,
                  lum_floor=None):
    """Engine inputs straight from a catalogue (no class constructor).

    ``zrange=(zmin, zmax)`` and ``lum_floor`` (a value, or a callable mapping this catalogue's own minimum to the one to
    use) pin the quadrature grid instead of deriving it from this catalogue's extremes: the shards of ONE catalogue held
    by several ranks must integrate over the same grid.

    Mirrors the reference's set-up arithmetic: knots ``zint = linspace(.95 zmin, 1.05 zmax, nknots)``
    (lumfuncmcmc.py:183, with a fixed knot count instead of N), ``lum`` from the exact D_L (:186, :259),
    the quadrature grid of ``setlnsimple`` with ``min_comp_frac = 0`` (:217-232).  With ``tabulated=True`` also
    returns ``Om_arr`` / ``integ_part`` evaluated *directly* (the reference tabulates them through a bicubic
    spline, :204-215, :233-235 -- these are engine inputs either way).
    """
    from scipy.interpolate import interp1d
    z = np.concatenate(cat['z'])
    flux = 1.0e-17 * np.concatenate(cat['flux'])
    K = len(cat['Flim'])
    fi = np.asarray(cat['field_ind'], dtype=np.int64)
    zmin, zmax = (float(z.min()), float(z.max())) if zrange is None else (float(zrange[0]), float(zrange[1]))
    zint = np.linspace(0.95 * zmin, 1.05 * zmax, nknots)
    DLarr = cosmo.luminosity_distance(zint)
    dVdzarr = cosmo.differential_comoving_volume(zint)
    DLf, dVdzf = interp1d(zint, DLarr), interp1d(zint, dVdzarr)
    DL = np.concatenate(cat['_DL']) if '_DL' in cat else _lumdist(z)       # the exact D_L of every source (:186)
    lum = np.log10(4.0 * np.pi * (DL * 3.086e24) ** 2 * flux)
    zarr = np.linspace(zmin, zmax, size_ln)
    floor = lum.min() if len(lum) else Lh - 6.0
    if lum_floor is not None:
        floor = float(lum_floor(floor)) if callable(lum_floor) else float(lum_floor)
    col = np.linspace(floor, Lh, size_ln)
    grid = np.repeat(col[:, None], size_ln, axis=1)
    inp = dict(lum=lum, z=z, zint=zint, DLarr=DLarr, dVdzarr=dVdzarr, field_ind=fi,
               Omega_0=np.asarray(cat['Omega_0'], dtype=np.float64), Flim=np.asarray(cat['Flim'], dtype=np.float64),
               alpha=float(cat['alpha']), fcmin=float(cat['fcmin']), logL=np.repeat(grid[None], K, axis=0),
               zarr=zarr, DL_zarr=DLf(zarr), volume_part=dVdzf(zarr), flux=flux, zmin=zmin, zmax=zmax,
               Lstar_lims=list(configLF.Lstar_lims), phistar_lims=list(configLF.phistar_lims),
               sch_al_lims=list(configLF.sch_al_lims), Flim_lims=list(configLF.Flim_lims),
               alpha_lims=list(configLF.alpha_lims), sch_al=configLF.sch_al, fix_sch_al=False,
               z1=z_pivots[0], z2=z_pivots[1], z3=z_pivots[2])
    if tabulated:
        om0_src = np.zeros(len(lum), dtype=int)
        flim_src = np.zeros(len(lum))
        for k in range(K):
            om0_src[fi[k]:fi[k + 1]] = inp['Omega_0'][k]
            flim_src[fi[k]:fi[k + 1]] = inp['Flim'][k]
        fsrc = 10 ** lum / (4.0 * np.pi * (3.086e24 * DLf(z)) ** 2)
        inp['Om_arr'] = om0_src / SQARCSEC * _fleming(fsrc, 1.0e-17 * flim_src, inp['alpha'], inp['fcmin'])
        zrep = np.repeat(zarr[None], size_ln, axis=0)
        fgrid = 10 ** grid / (4.0 * np.pi * (3.086e24 * DLf(zrep)) ** 2)
        inp['integ_part'] = np.stack([inp['volume_part'] * (inp['Omega_0'][k] / SQARCSEC) *
                                      _fleming(fgrid, 1.0e-17 * inp['Flim'][k], inp['alpha'], inp['fcmin'])
                                      for k in range(K)])
    return inp


def draw_thetas(inp, kind, W, seed=0, mode='near', scale=0.02, truth=None):
    """Walker positions: ``mode='prior'`` ~ U(prior box) like the reference's initialisation
    (lumfuncmcmc.py:436-446), ``mode='near'`` ~ truth + scale*N(0,1) (a converged ensemble)."""
    rng = np.random.default_rng(seed)
    K = len(inp['Flim'])
    if kind == 'z':
        lims = [inp['Lstar_lims']] * 3 + [inp['phistar_lims']] * 3
        centre = [42.5, 42.5, 42.5, -2.0, -2.0, -2.0]
        if not inp.get('fix_sch_al', False):
            lims, centre = lims + [inp['sch_al_lims']], centre + [-1.49]
    else:
        lims, centre = [inp['Lstar_lims'], inp['phistar_lims']], [42.5, -2.0]
        if not inp.get('fix_sch_al', False):
            lims, centre = lims + [inp['sch_al_lims']], centre + [-1.49]
        if kind == 'free':
            lims = lims + [inp['Flim_lims']] * K + [inp['alpha_lims']]
            centre = centre + list(inp['Flim']) + [inp['alpha']]
    lims = np.asarray(lims, dtype=np.float64)
    if truth is not None:
        centre = truth
    if mode == 'prior':
        return lims[:, 0] + rng.random((W, len(lims))) * (lims[:, 1] - lims[:, 0])
    return np.asarray(centre)[None, :] + scale * rng.standard_normal((W, len(lims)))
