"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy/SciPy) of the reference's hot path.

This file is the *checker*: only ``tests/`` (the pytest suites and the ``tests/run_*.py`` multi-GPU / end-to-end
check scripts), ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it.  The product (``lumfuncmcmc_b200``) never does, and
has no CPU fallback.

Each function restates one reference function on plain arrays (an "inputs" dict instead of ``self``), citing
the reference file:line it follows, and keeps the reference's order of floating-point operations so that it
reproduces the reference's values bit-for-bit (same NumPy ufuncs, same SciPy ``interp1d`` / ``trapezoid``).

Pinning: the reference ships no tests or golden vectors.  The oracle is pinned against
outputs of the reference's own code, imported unmodified behind stubs in the authoring container
(``oracle/refstub.py`` + ``oracle/make_golden.py`` -> ``tests/golden/*.npz``); ``tests/test_oracle_golden.py``
replays those fixtures.  Third-party arithmetic that is NOT under /root/reference: astropy's LambdaCDM
(version unpinned; restated in ``lumfuncmcmc_b200/cosmology.py``; "parity unpinned", feeds set-up tables only)
and emcee (unpinned; not on the likelihood path).

Inputs dict ("engine inputs", all float64 unless noted):
  lum[N], z[N]               per-source log10 luminosity and redshift, sources sorted by field
  zint[M], DLarr[M]          knots of the linear D_L(z) interpolant (Mpc)          (lumfuncmcmc.py:183-196)
  field_ind[K+1] (int64)     cumulative per-field offsets                          (lumfuncmcmc.py:148)
  Omega_0[K]                 field areas (arcsec^2, float); the per-source copy is int-truncated (:285)
  Flim[K], alpha, fcmin      configured completeness parameters (x1e-17 for Flim)
  logL[K,S,S], zarr[S], volume_part[S]    quadrature grid                            (lumfuncmcmc.py:217-232)
  Om_arr[N], integ_part[K,S,S]            tabulated-Omega variants                   (lumfuncmcmc.py:233-235)
  *_lims                      prior boxes;  fix_sch_al, sch_al (value used when fixed); z1,z2,z3 pivots
"""
import numpy as np
from scipy.integrate import trapezoid
from scipy.interpolate import interp1d

#: arcsec^2 per steradian (reference VmaxLumFunc.py:43)
SQARCSEC = (180. / np.pi * 3600.0) ** 2


# ------------------------------------------------------------------------------------------------
# L2 functions
# ------------------------------------------------------------------------------------------------
def schechter_log(logL, sch_al, logLstar, logphistar):
    """Schechter function per dex (reference lumfuncmcmc.py:25-44, twin lumfuncmcmc_z.py:69-88)."""
    dex = logL - logLstar
    return np.log(10.0) * 10 ** logphistar * 10 ** (dex * (sch_al + 1)) * np.exp(-10 ** dex)


def quad_coef(y1, y2, y3, z1, z2, z3):
    """Parabola through three pivots, reference operation order (reference lumfuncmcmc_z.py:26-43)."""
    a = ((y3 - y1) + (y2 - y1) * (z1 - z3) / (z2 - z1)) / (z3 ** 2 - z1 ** 2 + (z2 ** 2 - z1 ** 2) * (z1 - z3) / (z2 - z1))
    b = (y2 - y1 - a * (z2 ** 2 - z1 ** 2)) / (z2 - z1)
    c = y1 - a * z1 ** 2 - b * z1
    return a, b, c


def schechter_evolving(logL, z, sch_al, L123, phi123, z123):
    """Schechter with quadratic-in-z logL*, logphi* (reference lumfuncmcmc_z.py:45-67)."""
    aphi, bphi, cphi = quad_coef(phi123[0], phi123[1], phi123[2], *z123)
    alum, blum, clum = quad_coef(L123[0], L123[1], L123[2], *z123)
    phistar = aphi * z ** 2 + bphi * z + cphi
    Lstar = alum * z ** 2 + blum * z + clum
    return schechter_log(logL, sch_al, Lstar, phistar)


def f_tau(f50, alpha, fcmin):
    """Flux where the plain Fleming curve equals fcmin (reference VmaxLumFunc.py:143-167)."""
    a = (2 * fcmin - 1) ** 2.
    b = -1 * (abs(a / (1 - a)) * alpha ** -2.) ** 0.5
    return f50 * 10 ** b


def fleming(f, f50, alpha, fcmin):
    """(Modified) Fleming completeness (reference VmaxLumFunc.py:95-127, 136-141)."""
    num = alpha * np.log10(f / f50)
    den = (1. + num ** 2.) ** 0.5
    fc = 0.5 * (1. + num / den)
    if not fcmin:
        return fc
    decay = 1. - np.exp(-f / f_tau(f50, alpha, fcmin))
    return fc ** (1. / decay)


def omega(logL, z, DLf, Omega_0, f50, alpha, fcmin):
    """Effective solid angle [sr] at (logL, z) (reference lumfuncmcmc.py:47-70)."""
    L = 10 ** logL
    return Omega_0 / SQARCSEC * fleming(L / (4.0 * np.pi * (3.086e24 * DLf(z)) ** 2), f50, alpha, fcmin)


def _trapz2(integ, logL, zarr):
    """trapz over luminosity (axis 0, 2-D abscissa) then over redshift (reference lumfuncmcmc.py:377)."""
    return trapezoid(trapezoid(integ, logL, axis=0), zarr)


# ------------------------------------------------------------------------------------------------
# model state (what the reference keeps on ``self``)
# ------------------------------------------------------------------------------------------------
class FreeCompModel:
    """Single-z model; ``fix_comp`` selects lnlike_fix_comp (reference lumfuncmcmc.py:320-424)."""

    def __init__(self, inp, fix_comp=False):
        self.inp = inp
        self.fix_comp = bool(fix_comp)
        self.K = len(inp['Flim'])
        self.fix_sch_al = bool(inp.get('fix_sch_al', False))
        self.DLf = interp1d(inp['zint'], inp['DLarr'])
        fi = np.asarray(inp['field_ind'], dtype=np.int64)
        n = int(fi[-1])
        self.field_ind = fi
        # per-source copies; the area copy is integer-typed (reference lumfuncmcmc.py:283-288)
        self.Flims_arr = np.zeros(n)
        self.Omega_0_arr = np.zeros(n, dtype=int)
        for k in range(self.K):
            self.Omega_0_arr[fi[k]:fi[k + 1]] = inp['Omega_0'][k]
        self.zarr_rep = np.repeat(np.asarray(inp['zarr'])[None], len(inp['zarr']), axis=0)
        self.sch_al = inp['sch_al']
        self.Flim = list(inp['Flim'])
        self.alpha = inp['alpha']
        self.Lstar = self.phistar = None

    # reference lumfuncmcmc.py:320-337
    def unpack(self, theta):
        K = self.K
        self.Lstar, self.phistar = theta[0], theta[1]
        if self.fix_comp:
            if not self.fix_sch_al:
                self.sch_al = theta[2]
        elif self.fix_sch_al:
            self.Flim, self.alpha = theta[2:2 + K], theta[2 + K]
        else:
            self.sch_al = theta[2]
            self.Flim, self.alpha = theta[3:3 + K], theta[3 + K]

    # reference lumfuncmcmc.py:339-358 (inclusive box; NaN compares False everywhere)
    def lnprior(self):
        inp = self.inp
        flag = 1.0
        for val, lims in ((self.Lstar, inp['Lstar_lims']), (self.phistar, inp['phistar_lims']),
                          (self.sch_al, inp['sch_al_lims'])):
            flag *= ((val >= lims[0]) * (val <= lims[1]))
        for k in range(self.K):
            flag *= ((self.Flim[k] >= inp['Flim_lims'][0]) * (self.Flim[k] <= inp['Flim_lims'][1]))
        flag *= ((self.alpha >= inp['alpha_lims'][0]) * (self.alpha <= inp['alpha_lims'][1]))
        return 0.0 if flag else -np.inf

    # reference lumfuncmcmc.py:360-378
    def lnlike(self):
        inp = self.inp
        fi = self.field_ind
        for k in range(self.K):
            self.Flims_arr[fi[k]:fi[k + 1]] = self.Flim[k]
        with np.errstate(divide='ignore', over='ignore', under='ignore', invalid='ignore'):
            lnpart = np.log(schechter_log(inp['lum'], self.sch_al, self.Lstar, self.phistar) *
                            omega(inp['lum'], inp['z'], self.DLf, self.Omega_0_arr, 1.0e-17 * self.Flims_arr,
                                  self.alpha, inp['fcmin'])).sum()
            fullint = 0.0
            for k in range(self.K):
                integ_part = inp['volume_part'] * omega(inp['logL'][k], self.zarr_rep, self.DLf, inp['Omega_0'][k],
                                                        1.0e-17 * self.Flim[k], self.alpha, inp['fcmin'])
                integ = schechter_log(inp['logL'][k], self.sch_al, self.Lstar, self.phistar) * integ_part
                fullint += _trapz2(integ, inp['logL'][k], inp['zarr'])
        return lnpart - fullint

    # reference lumfuncmcmc.py:380-393
    def lnlike_fix_comp(self):
        inp = self.inp
        with np.errstate(divide='ignore', over='ignore', under='ignore', invalid='ignore'):
            lnpart = np.log(schechter_log(inp['lum'], self.sch_al, self.Lstar, self.phistar) * inp['Om_arr']).sum()
            fullint = 0.0
            for k in range(self.K):
                integ = schechter_log(inp['logL'][k], self.sch_al, self.Lstar, self.phistar) * inp['integ_part'][k]
                fullint += _trapz2(integ, inp['logL'][k], inp['zarr'])
        return lnpart - fullint

    # reference lumfuncmcmc.py:395-424
    def lnprob(self, theta):
        self.unpack(theta)
        lp = self.lnprior()
        if not np.isfinite(lp):
            return -np.inf
        return (self.lnlike_fix_comp() if self.fix_comp else self.lnlike()) + lp


class EvolvingModel:
    """Redshift-evolving model (reference lumfuncmcmc_z.py:332-392)."""

    def __init__(self, inp):
        self.inp = inp
        self.K = len(inp['Flim'])
        self.fix_sch_al = bool(inp.get('fix_sch_al', False))
        self.sch_al = inp['sch_al']
        self.z123 = (inp['z1'], inp['z2'], inp['z3'])
        self.zarr_rep = np.repeat(np.asarray(inp['zarr'])[None], len(inp['zarr']), axis=0)
        self.L123 = self.phi123 = None

    def unpack(self, theta):                                    # lumfuncmcmc_z.py:332-341
        self.L123 = (theta[0], theta[1], theta[2])
        self.phi123 = (theta[3], theta[4], theta[5])
        if not self.fix_sch_al:
            self.sch_al = theta[6]

    def lnprior(self):                                          # lumfuncmcmc_z.py:343-362 (strict on L, phi)
        inp = self.inp
        if self.fix_sch_al:
            flag = 1
        else:
            flag = ((self.sch_al >= inp['sch_al_lims'][0]) * (self.sch_al <= inp['sch_al_lims'][1]))
        for i in range(3):
            flag *= ((self.L123[i] > inp['Lstar_lims'][0]) * (self.L123[i] < inp['Lstar_lims'][1]))
            flag *= ((self.phi123[i] > inp['phistar_lims'][0]) * (self.phi123[i] < inp['phistar_lims'][1]))
        return 0.0 if flag else -np.inf

    def lnlike(self):                                           # lumfuncmcmc_z.py:364-376
        inp = self.inp
        with np.errstate(divide='ignore', over='ignore', under='ignore', invalid='ignore'):
            lnpart = np.log(schechter_evolving(inp['lum'], inp['z'], self.sch_al, self.L123, self.phi123, self.z123) *
                            inp['Om_arr']).sum()
            fullint = 0.0
            for k in range(self.K):
                integ = schechter_evolving(inp['logL'][k], self.zarr_rep, self.sch_al, self.L123, self.phi123,
                                           self.z123) * inp['integ_part'][k]
                fullint += _trapz2(integ, inp['logL'][k], inp['zarr'])
        return lnpart - fullint

    def lnprob(self, theta):                                    # lumfuncmcmc_z.py:378-392
        self.unpack(theta)
        lp = self.lnprior()
        if not np.isfinite(lp):
            return -np.inf
        return self.lnlike() + lp


def make_model(inp, kind):
    """kind in {'free', 'fixed', 'z'} -> object with ``lnprob(theta)``."""
    if kind == 'free':
        return FreeCompModel(inp, fix_comp=False)
    if kind == 'fixed':
        return FreeCompModel(inp, fix_comp=True)
    if kind == 'z':
        return EvolvingModel(inp)
    raise ValueError(kind)


def lnprob_batch(inp, kind, thetas):
    """Serial loop over walkers, one ``lnprob`` per row -- how the reference is driven (lumfuncmcmc.py:489-491)."""
    model = make_model(inp, kind)
    thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    return np.array([model.lnprob(t) for t in thetas], dtype=np.float64)


# ------------------------------------------------------------------------------------------------
# 1/V_eff estimator and binned LF
# ------------------------------------------------------------------------------------------------
def veff_weights(flux, flims_arr, alpha, fcmin, sum_omega, vol_int, zmin, zmaxval=None):
    """phi_i = 1 / integral of (sum Omega_0/sqarcsec * fleming(F_i) * dV/dz) dz.

    Restates reference lumfuncmcmc.py:515-524 + VmaxLumFunc.py:215-257 for the case the integrand's
    flux factor is constant in z, so the z integral is the shared ``vol_int`` (or ``vol_int[i]`` per source
    when ``zmaxval`` varies); SURVEY.md section 3.4 verified this equals the reference's QUADPACK value to
    4e-16 when ``vol_int`` is itself QUADPACK's integral of the dV/dz interpolant.
    Sources whose upper limit does not exceed zmin keep phi = 0 (lumfuncmcmc.py:524).
    """
    comp = fleming(flux, 1.0e-17 * flims_arr, alpha, fcmin)
    phi = 1.0 / (sum_omega / SQARCSEC * comp * vol_int)
    if zmaxval is not None:
        phi = np.where(np.asarray(zmaxval) > zmin, phi, 0.0)
    return phi


def veff_volumes_loop(lum, fmin_per_source, zmin, zmax, dVdzf, luminosity_distance_cm):
    """Per-source upper redshift limit and volume exactly as the reference's loop does it (lumfuncmcmc.py:521-524):
    ``zmaxval = min(zmax, fsolve(4 pi D_L(z)^2 Fmin - L, 1.5))`` (VmaxLumFunc.py:722-753) and
    ``quad(dVdzf, zmin, zmaxval)`` (the z-dependent factor of VmaxLumFunc.py:230-232, :255).  O(N) SciPy calls: small N only.
    Returns (zmaxval, vol, valid); vol = 1 where not valid."""
    from scipy.integrate import quad
    from scipy.optimize import fsolve
    n = len(lum)
    zm, vol, valid = np.zeros(n), np.ones(n), np.zeros(n, dtype=bool)
    for i in range(n):
        L = 10 ** lum[i]
        root = fsolve(lambda x: 4.0 * np.pi * luminosity_distance_cm(x) ** 2 * fmin_per_source[i] - L, 1.5)[0]
        zm[i] = min(zmax, root)
        if zm[i] > zmin:
            vol[i] = quad(dVdzf, zmin, zm[i])[0]
            valid[i] = True
    return zm, vol, valid


def veff_volumes_vectorised(lum, fmin_per_source, zmin, zmax, zk, dVk, luminosity_distance_mpc, mpc_cm=3.085677581491367e24):
    """The same quantities without the per-source SciPy calls, for catalogue sizes the loop cannot reach: the root of the
    monotone D_L(z) by bisection to machine precision (what fsolve approximates to 1.5e-8) and the EXACT integral of the
    piecewise-linear interpolant through (zk, dVk) (what QUADPACK approximates to 1.5e-8)."""
    lum, fmin = np.asarray(lum, dtype=np.float64), np.asarray(fmin_per_source, dtype=np.float64)
    target = np.sqrt(10 ** lum / (4.0 * np.pi * fmin)) / mpc_cm                      # D_L [Mpc] at the flux limit
    dl_lo, dl_hi = float(luminosity_distance_mpc(zmin)), float(luminosity_distance_mpc(zmax))
    lo, hi = np.full(lum.shape, float(zmin)), np.full(lum.shape, float(zmax))
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        below = luminosity_distance_mpc(mid) < target
        lo, hi = np.where(below, mid, lo), np.where(below, hi, mid)
    zm = np.where(target >= dl_hi, zmax, np.where(target <= dl_lo, zmin, 0.5 * (lo + hi)))
    valid = zm > zmin
    seg = 0.5 * (dVk[1:] + dVk[:-1]) * np.diff(zk)
    cum = np.concatenate([[0.0], np.cumsum(seg)])

    def integral_to(z):
        j = np.clip(np.searchsorted(zk, z, side='right') - 1, 0, len(zk) - 2)
        dz = z - zk[j]
        slope = (dVk[j + 1] - dVk[j]) / (zk[j + 1] - zk[j])
        return cum[j] + dz * (dVk[j] + 0.5 * slope * dz)

    vol = np.where(valid, integral_to(zm) - integral_to(np.float64(zmin)), 1.0)
    return zm, vol, valid


def binned_lf_counts(L, edges):
    """Integer source counts in the half-open bins [e_j, e_{j+1}) (reference VmaxLumFunc.py:345-349)."""
    return np.array([np.count_nonzero(np.logical_and(L >= edges[j], L < edges[j + 1]))
                     for j in range(len(edges) - 1)], dtype=np.int64)


def boot_err_log(L, phi, nboot=100, nbin=25, Larr=None, rng_randint=None):
    """Binned LF + bootstrap variance (reference VmaxLumFunc.py:304-364, ``correct_low=False``).

    ``rng_randint(n, size)`` defaults to NumPy's legacy global ``np.random.randint`` as in the reference (:353).
    Returns (Lavg, lfbinorig, var, counts, Larr).
    """
    if rng_randint is None:
        rng_randint = lambda n, size: np.random.randint(n, size=size)
    if Larr is None:
        Larr = np.linspace(min(L) * 1.001, max(L), nbin + 1)
    nb = len(Larr) - 1
    Lavg = np.linspace((Larr[0] + Larr[1]) / 2.0, (Larr[-1] + Larr[-2]) / 2.0, nb)
    dL = Lavg[1] - Lavg[0]
    lfbinorig = np.zeros(nb)
    counts = np.zeros(nb, dtype=np.int64)
    for j in range(nb):
        sel = np.logical_and(L >= Larr[j], L < Larr[j + 1])
        counts[j] = np.count_nonzero(sel)
        if counts[j]:
            lfbinorig[j] = sum(phi[sel]) / dL
    lfbin = np.zeros((nboot, nb))
    for k in range(nboot):
        boot = rng_randint(len(phi), len(phi))
        Lb, pb = L[boot], phi[boot]
        for j in range(nb):
            sel = np.logical_and(Lb >= Larr[j], Lb < Larr[j + 1])
            if np.count_nonzero(sel):
                lfbin[k, j] = sum(pb[sel]) / dL
    binavg = np.average(lfbin, axis=0)
    var = 1. / (nboot - 1) * np.sum((lfbin - binavg) ** 2, axis=0)
    var[var <= 0.0] = min(var[var > 0.0])
    return Lavg, lfbinorig, var, counts, Larr
