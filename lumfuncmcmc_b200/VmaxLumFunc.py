"""1/V_eff luminosity-function estimator and completeness curve -- the public names of the reference's
VmaxLumFunc module that the MCMC classes and the drivers use:
``cosmo``, ``sqarcsec``, ``fleming``, ``expdecay``, ``inverse_fleming``, ``schechter_log``, ``lumfuncintv2``,
``lumfunc``, ``getBootErrLog``, ``get_L_constF``, ``getMaxz``.

The scalar helpers (completeness curve, single-source weight) are host-side NumPy: they are set-up / plotting
utilities.  The O(N) work -- per-source weights, luminosity binning, bootstrap replicates -- runs on the GPU
through the engine's C ABI (``lf_veff_bin`` / ``lf_bin_weights`` / ``lf_boot_bin``); there is no CPU fallback for it.
"""
import numpy as np
from scipy.integrate import quad
from scipy.optimize import fsolve

from .cosmology import cosmo as _cosmo

#: astropy-like view of the cosmology (``.luminosity_distance(z).value`` ...), reference VmaxLumFunc.py:16-17
cosmo = _cosmo.as_astropy_like()
#: arcsec^2 per steradian, reference VmaxLumFunc.py:43
sqarcsec = (180. / np.pi * 3600.0) ** 2


def schechter_log(L, al, phistar, Lstar):
    """Schechter function per dex with log10 parameters (reference VmaxLumFunc.py:54-56)."""
    x = L - Lstar
    return np.log(10.0) * 10 ** phistar * 10 ** (x * (al + 1)) * np.exp(-10 ** x)


def expdecay(x, tau):
    """1 - exp(-x/tau): the exponent of the faint-end modification (reference VmaxLumFunc.py:136-141)."""
    return 1. - np.exp(-x / tau)


def inverse_fleming(f50, alpha, fcmin=0.1):
    """Flux at which the plain Fleming curve equals ``fcmin`` (reference VmaxLumFunc.py:143-167)."""
    a = (2 * fcmin - 1) ** 2.
    return f50 * 10 ** (-1 * (abs(a / (1 - a)) * alpha ** -2.) ** 0.5)


def fleming(f, Flim=3.0e-17, alpha=3.5, fcmin=0.1):
    """Fleming completeness fraction at linear flux ``f``; with ``fcmin`` truthy the curve is raised to the power
    1/(1 - exp(-f/f_tau)) so that it falls to zero faster below f_tau (reference VmaxLumFunc.py:95-127).
    ``alpha=None`` means "no completeness correction" and returns ones."""
    if alpha is None:
        return np.ones(len(list(f)))
    slope = alpha * np.log10(f / Flim)
    fc = 0.5 * (1. + slope / (1. + slope ** 2.) ** 0.5)
    if not fcmin:
        return fc
    return fc ** (1. / expdecay(f, inverse_fleming(f50=Flim, alpha=alpha, fcmin=fcmin)))


def lumfuncintv2(z, F, Omega_0, func, Flim, alpha, fcmin=0.1):
    """Integrand of the effective volume: solid angle [sr] x completeness x dV/dz/dOmega (reference
    VmaxLumFunc.py:215-232).  The flux does not depend on z here, so this is a constant times ``func(z)``."""
    return Omega_0 / sqarcsec * fleming(F, Flim, alpha, fcmin=fcmin) * func(z)


def lumfunc(F, func, Omega_0=100.0, minz=1.16, maxz=1.9, Flim=3.0e-17, alpha=3.5, fcmin=0.1):
    """1/V_eff weight of ONE source (reference VmaxLumFunc.py:235-257).  Kept for API compatibility; whole
    catalogues go through :func:`veff_weights_gpu`."""
    vol, _ = quad(lumfuncintv2, minz, maxz, args=(F, Omega_0, func, Flim, alpha, fcmin))
    return 1.0 / vol


def get_L_constF(F, z):
    """Luminosity [erg/s] of flux F at redshift z (reference VmaxLumFunc.py:722-737)."""
    return 4.0 * np.pi * (cosmo.luminosity_distance(z).to('cm').value) ** 2 * F


def getMaxz(L, Fmin):
    """Redshift at which luminosity L drops to flux Fmin (reference VmaxLumFunc.py:739-753)."""
    return fsolve(lambda x: get_L_constF(Fmin, x) - L, 1.5)[0]


# --------------------------------------------------------------------------------------------------
# GPU paths
# --------------------------------------------------------------------------------------------------
_engine_cache = {}


def _veff_engine(device=0):
    from .engine import VeffEngine
    if device not in _engine_cache:
        _engine_cache[device] = VeffEngine(device=device)
    return _engine_cache[device]


def veff_weights_gpu(flux, lum, field_ind, Flim, alpha, fcmin, sum_omega, vol_int, edges, vol_per_source=None,
                     valid=None, device=0, engine=None):
    """Per-source 1/V_eff weights and the binned LF of the sample in one streaming GPU pass.

    phi_i = valid_i / (sum_omega/sqarcsec * fleming(flux_i; 1e-17 Flim[field(i)], alpha, fcmin) * vol_i)
    (reference lumfuncmcmc.py:521-524 with VmaxLumFunc.py:256-257).  Returns (phi, counts, sumphi)."""
    eng = engine or _veff_engine(device)
    return eng.veff_bin(flux, lum, field_ind, Flim, alpha, fcmin, sum_omega, vol_int, edges,
                        vol_per_source=vol_per_source, valid=valid)


def getBootErrLog(L, phi, minz, maxz, nboot=100, nbin=25, Fmin=1.0e-20, Larr=None, correct_low=False, device=0,
                  engine=None, return_counts=False, rng='auto', seed=None, Lrange=None):
    """Binned luminosity function dn/dlogL with bootstrap variances (reference VmaxLumFunc.py:304-364).

    Bin edges ``linspace(min(L)*1.001, max(L), nbin+1)`` unless ``Larr`` is given; bins are half-open
    [e_j, e_{j+1}); the bootstrap draws ``np.random.randint(N, size=N)`` from NumPy's global stream exactly as the
    reference does (:353), so a seeded run resamples the same sources.  The index draw happens on the host (it is
    the reference's RNG); the gather + binning of every replicate is one GPU pass over per-source multiplicities.
    ``correct_low`` (partial-bin correction, never enabled by the MCMC classes) is not supported.

    ``rng='mt19937'`` draws the SAME stream on the GPU: NumPy's global MT19937 state is handed to the device
    (``lf_boot_mt_set_state``), every replicate's indices are generated there exactly as ``np.random.randint(N, size=N)``
    would (masked rejection sampling of 32-bit outputs), and the host generator is re-synchronised afterwards -- bit-identical
    replicates and generator state, without N host draws and an N-element upload per replicate.  ``rng='auto'`` (default)
    takes that route from 1e5 sources up and the host draw below; the results do not depend on the choice.

    ``rng='device'`` (additive option) resamples on the GPU instead: a Philox stream keyed by ``seed`` (default: one draw
    from NumPy's global stream) generates every replicate's indices and multiplicities on the device -- statistically
    equivalent variances, 100 replicates of 1e7 sources in a fraction of a second instead of half a minute of host RNG.

    ``phi=None`` (with ``engine``): the sample and its weights are the ones ``engine`` already holds on the device (left
    there by ``veff_bin_resident``) -- nothing per-source is uploaded.  ``Lrange=(min(L), max(L))`` spares the two passes
    over ``L`` when the caller knows them.
    """
    if correct_low:
        raise NotImplementedError("correct_low=True is outside the supported path (the MCMC classes never pass it)")
    L = np.ascontiguousarray(L, dtype=np.float64)
    resident = phi is None
    if resident and engine is None:
        raise ValueError("phi=None needs the engine that holds the weights")
    if not resident:
        phi = np.ascontiguousarray(phi, dtype=np.float64)
    if Larr is None:
        print("Min Luminosity:", np.log10(get_L_constF(Fmin, maxz)))
        lo, hi = Lrange if Lrange is not None else (np.min(L), np.max(L))
        Larr = np.linspace(lo * 1.001, hi, nbin + 1)
    Larr = np.ascontiguousarray(Larr, dtype=np.float64)
    nb = len(Larr) - 1
    Lavg = np.linspace((Larr[0] + Larr[1]) / 2.0, (Larr[-1] + Larr[-2]) / 2.0, nb)
    dL = Lavg[1] - Lavg[0]
    eng = engine or _veff_engine(device)
    counts, sums = eng.bin_weights(None, None, Larr) if resident else eng.bin_weights(L, phi, Larr)
    lfbinorig = np.where(counts > 0, sums / dL, 0.0)
    lfbin = np.zeros((nboot, nb))
    n = len(L)
    if rng not in ('auto', 'host', 'mt19937', 'device'):
        raise ValueError("rng must be 'auto', 'host', 'mt19937' or 'device'")
    mt_ok = 2 <= n < 2 ** 32 and np.random.get_state()[0] == 'MT19937'
    if rng == 'auto':
        rng = 'mt19937' if (n >= 100000 and mt_ok) else 'host'
    if rng == 'mt19937':
        if not mt_ok:
            raise ValueError("rng='mt19937' needs 2 <= N < 2**32 sources and NumPy's legacy MT19937 global generator")
        eng.boot_mt_set_state(np.random.get_state())
    if rng == 'device' and seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 32)
    for k in range(nboot):
        if rng == 'device':
            bc, bs = eng.boot_bin_device(seed, k)
        elif rng == 'mt19937':
            bc, bs = eng.boot_bin_mt()
        else:
            boot = np.random.randint(n, size=n)
            bc, bs = eng.boot_bin(np.bincount(boot, minlength=n))
        lfbin[k] = np.where(bc > 0, bs / dL, 0.0)
    if rng == 'mt19937':
        np.random.set_state(eng.boot_mt_get_state())       # the host generator continues where the reference's would
    binavg = np.average(lfbin, axis=0)
    var = 1. / (nboot - 1) * np.sum((lfbin - binavg) ** 2, axis=0)
    var[var <= 0.0] = min(var[var > 0.0])
    if return_counts:
        return Lavg, lfbinorig, var, counts
    return Lavg, lfbinorig, var
