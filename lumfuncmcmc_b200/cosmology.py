"""Host-side FLRW distances for the set-up tables (luminosity distance, dV/dz/dOmega).

The reference builds its tables from ``astropy.cosmology.LambdaCDM(H0=70, Tcmb0=2.725,
Om0=0.3, Ode0=0.7)`` (reference VmaxLumFunc.py:14-17; call sites lumfuncmcmc.py:186-188,
lumfuncmcmc_z.py:230-232, VmaxLumFunc.py:737).  astropy is not in this image and its version is not
pinned by the reference, so this module restates the published LambdaCDM arithmetic (photons + 3.04
massless neutrinos from Tcmb0, curvature from closure) -- "parity unpinned" for the table VALUES, which
is harmless for the hot path because the tables are inputs to both the oracle and the engine.

Everything here runs once per fit, on the host, vectorised: a panel-wise Gauss-Legendre cumulative
integral of 1/E(z) plus one 8-point Gauss-Legendre closure per query point (O(N) instead of the
reference's per-point QUADPACK call).
"""
import numpy as np

_C_KM_S = 299792.458                 # km/s
_SIGMA_SB = 5.670374419e-8           # W m^-2 K^-4
_G = 6.6743e-11                      # m^3 kg^-1 s^-2
_C_M_S = 299792458.0
_MPC_M = 3.085677581491367e22
_MPC_CM = 3.085677581491367e24

_GL_X, _GL_W = np.polynomial.legendre.leggauss(8)


class _Quantity:
    """Minimal stand-in for an astropy Quantity: ``.value`` and ``.to('cm').value``."""

    def __init__(self, value, unit):
        self.value = value
        self.unit = unit

    def to(self, unit):
        if unit == self.unit:
            return _Quantity(self.value, unit)
        if self.unit == 'Mpc' and unit == 'cm':
            return _Quantity(self.value * _MPC_CM, 'cm')
        raise ValueError("unsupported conversion %s -> %s" % (self.unit, unit))


class LambdaCDM:
    """FLRW cosmology with a cosmological constant and (possibly) curvature.

    ``luminosity_distance(z)`` [Mpc] and ``differential_comoving_volume(z)`` [Mpc^3 / sr] return plain
    float64 arrays (scalars for scalar input).  ``as_astropy_like()`` wraps them in objects carrying
    ``.value`` for code written against astropy.
    """

    def __init__(self, H0=70.0, Om0=0.3, Ode0=0.7, Tcmb0=2.725, Neff=3.04, panel=0.01):
        self.H0 = float(getattr(H0, 'value', H0))
        self.Om0, self.Ode0 = float(Om0), float(Ode0)
        self.Tcmb0 = float(getattr(Tcmb0, 'value', Tcmb0))
        self.Neff = float(Neff)
        h0_si = self.H0 * 1.0e3 / _MPC_M
        rho_crit = 3.0 * h0_si ** 2 / (8.0 * np.pi * _G)              # kg m^-3
        rho_gamma = 4.0 * _SIGMA_SB / _C_M_S ** 3 * self.Tcmb0 ** 4   # kg m^-3
        self.Ogamma0 = rho_gamma / rho_crit
        self.Onu0 = 0.22710731766 * self.Neff * self.Ogamma0           # 7/8 (4/11)^(4/3)
        self.Or0 = self.Ogamma0 + self.Onu0
        self.Ok0 = 1.0 - self.Om0 - self.Ode0 - self.Or0
        self.hubble_distance = _C_KM_S / self.H0                       # Mpc
        self._panel = float(panel)
        self._cum = np.zeros(1)                                        # cumulative int_0^{p*panel} dz/E

    # -- E(z) ----------------------------------------------------------------------------------
    def efunc(self, z):
        zp1 = 1.0 + np.asarray(z, dtype=np.float64)
        return np.sqrt(zp1 * zp1 * ((self.Or0 * zp1 + self.Om0) * zp1 + self.Ok0) + self.Ode0)

    def _extend(self, zmax):
        need = int(np.ceil(zmax / self._panel)) + 1
        have = len(self._cum) - 1
        if need <= have:
            return
        lo = self._panel * np.arange(have, need)
        half = 0.5 * self._panel
        nodes = lo[:, None] + half * (1.0 + _GL_X[None, :])
        panels = half * (1.0 / self.efunc(nodes)) @ _GL_W
        self._cum = np.concatenate([self._cum, self._cum[-1] + np.cumsum(panels)])

    def _dc_over_dh(self, z):
        z = np.asarray(z, dtype=np.float64)
        if z.size == 0:
            return np.zeros_like(z)
        if np.any(z < 0):
            raise ValueError("negative redshift")
        self._extend(float(np.max(z)))
        p = np.floor(z / self._panel).astype(np.int64)
        lo = p * self._panel
        half = 0.5 * (z - lo)
        acc = np.zeros_like(z)
        for x, w in zip(_GL_X, _GL_W):
            acc += w / self.efunc(lo + half * (1.0 + x))
        return self._cum[p] + half * acc

    # -- distances -----------------------------------------------------------------------------
    def comoving_transverse_distance(self, z):
        dc = self._dc_over_dh(z)
        ok = self.Ok0
        if ok == 0.0:
            dm = dc
        elif ok > 0.0:
            s = np.sqrt(ok)
            dm = np.sinh(s * dc) / s
        else:
            s = np.sqrt(-ok)
            dm = np.sin(s * dc) / s
        return self.hubble_distance * dm

    def luminosity_distance(self, z):
        z = np.asarray(z, dtype=np.float64)
        return (1.0 + z) * self.comoving_transverse_distance(z)

    def differential_comoving_volume(self, z):
        dm = self.comoving_transverse_distance(z)
        return self.hubble_distance * dm * dm / self.efunc(z)

    def as_astropy_like(self):
        return _AstropyLike(self)


class _AstropyLike:
    """Presents the astropy method signatures the reference's code uses (values in ``.value``)."""

    def __init__(self, cosmo):
        self._c = cosmo

    def luminosity_distance(self, z):
        return _Quantity(self._c.luminosity_distance(z), 'Mpc')

    def differential_comoving_volume(self, z):
        return _Quantity(self._c.differential_comoving_volume(z), 'Mpc3/sr')


#: the cosmology object the reference instantiates at import (VmaxLumFunc.py:16-17)
cosmo = LambdaCDM(H0=70.0, Tcmb0=2.725, Om0=0.3, Ode0=0.7)
