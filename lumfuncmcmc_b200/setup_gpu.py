"""Set-up tables on the GPU: the O(N) pieces of the reference's constructor chain.

The reference builds its distance tables on N knots (N = number of sources) and evaluates the resulting linear
interpolants and the cosmology once per source (reference lumfuncmcmc.py:180-202, :69-70, :235).  On the host these
are single-threaded passes with a cache-missing binary search per point (17 s per interpolant at N = 1e7); here
they are two kernels behind the C ABI:

* :func:`interp_linear` -- ``numpy.interp`` / ``scipy.interpolate.interp1d(kind='linear')`` BIT FOR BIT
  (``lf_interp_linear``);
* :func:`cosmo_distances` -- luminosity distance and dV/dz/dOmega with the arithmetic of
  :mod:`lumfuncmcmc_b200.cosmology`, operation for operation (``lf_cosmo_distances``).

:class:`LinearTable` is the drop-in for the ``interp1d`` objects the classes keep (``DLf``, ``dVdzf``, ``minlumf``):
large evaluations go to the GPU when one is present, small ones (grids of a few hundred points) stay on the host,
where NumPy evaluates the same arithmetic.  Without a GPU everything stays on the host: these are set-up tables,
not the likelihood path.
"""
import ctypes as C

import numpy as np

from . import _lib

#: evaluations at least this long go to the GPU (below it the PCIe round trip costs more than the host pass)
GPU_MIN_POINTS = 200000

_ndev = None


def gpu_count():
    global _ndev
    if _ndev is None:
        try:
            _ndev = int(_lib.load().lf_device_count())
        except Exception:
            _ndev = 0
    return _ndev


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def interp_linear(xk, yk, x, device=0):
    """``numpy.interp(x, xk, yk)`` on the GPU, bit-identical; ValueError outside the knot range (as interp1d)."""
    lib = _lib.load()
    xk, yk, x = _f64(xk), _f64(yk), _f64(x)
    y = np.empty_like(x)
    rc = lib.lf_interp_linear(int(device), xk.shape[0], _p(xk), _p(yk), x.size, _p(x), _p(y))
    if rc:
        msg = lib.lf_last_error().decode()
        if 'outside the interpolation range' in msg:
            raise ValueError("A value in x_new is outside the interpolation range.")
        raise _lib.EngineError(msg)
    return y


def cosmology_struct(cosmo, zmax):
    """(``lf_cosmology`` struct, cumulative panel integrals covering [0, zmax]) of a :class:`..cosmology.LambdaCDM`."""
    cosmo._extend(float(zmax))
    cum = _f64(cosmo._cum)
    c = _lib.LFCosmology()
    c.H0, c.Om0, c.Ode0, c.Or0, c.Ok0, c.panel = cosmo.H0, cosmo.Om0, cosmo.Ode0, cosmo.Or0, cosmo.Ok0, cosmo._panel
    glx, glw = np.polynomial.legendre.leggauss(8)
    c.gl_x = (C.c_double * 8)(*glx)
    c.gl_w = (C.c_double * 8)(*glw)
    return c, cum


def cosmo_distances(cosmo, z, device=0, want_dl=True, want_dv=True):
    """(D_L [Mpc], dV/dz/dOmega [Mpc^3/sr]) of a :class:`lumfuncmcmc_b200.cosmology.LambdaCDM` for an array z."""
    lib = _lib.load()
    z = _f64(z)
    shape = z.shape
    z = z.ravel()
    if z.size == 0:
        return (np.zeros(shape) if want_dl else None), (np.zeros(shape) if want_dv else None)
    c, cum = cosmology_struct(cosmo, float(np.max(z)))
    dl = np.empty_like(z) if want_dl else None
    dv = np.empty_like(z) if want_dv else None
    _lib.check(lib.lf_cosmo_distances(int(device), C.byref(c), _p(cum), cum.shape[0], z.size, _p(z), _p(dl), _p(dv)), lib)
    return (dl.reshape(shape) if want_dl else None), (dv.reshape(shape) if want_dv else None)


def omega_sources(lum, z, field_ind, omega0_int, flim, alpha, fcmin, DLf, device=0):
    """Per-source tabulated Omega on the GPU (``lf_omega_sources``): what the reference evaluates with
    ``Omega(lum, z, DLf, Omega_0_arr, 1e-17 * Flims_arr, alpha, fcmin)`` (lumfuncmcmc.py:235)."""
    lib = _lib.load()
    lum, z = _f64(lum), _f64(z)
    fi = np.ascontiguousarray(field_ind, dtype=np.int64)
    om0 = np.ascontiguousarray(omega0_int, dtype=np.int64)
    flim = _f64(flim)
    out = np.empty_like(lum)
    rc = lib.lf_omega_sources(int(device), lum.shape[0], _p(lum), _p(z), _p(fi), len(flim), _p(om0), _p(flim), float(alpha),
                              float(fcmin) if fcmin else 0.0, DLf.x.shape[0], _p(DLf.x), _p(DLf.y), _p(out))
    if rc:
        msg = lib.lf_last_error().decode()
        if 'outside the interpolation range' in msg:
            raise ValueError("A value in x_new is outside the interpolation range.")
        raise _lib.EngineError(msg)
    return out


class LinearTable:
    """Linear interpolant with ``interp1d``'s call semantics (bounds error, array in / array out) and NumPy's
    arithmetic; ``.x`` / ``.y`` are the knots."""

    def __init__(self, x, y, device=0):
        self.x, self.y = _f64(x), _f64(y)
        if self.x.ndim != 1 or self.x.shape != self.y.shape or self.x.shape[0] < 2:
            raise ValueError("x and y must be 1-D arrays of equal length >= 2")
        self.device = device

    def __call__(self, x_new):
        x_new = np.asarray(x_new, dtype=np.float64)
        scalar = x_new.ndim == 0
        flat = np.ascontiguousarray(x_new.ravel())
        if flat.size >= GPU_MIN_POINTS and gpu_count() > 0:
            out = interp_linear(self.x, self.y, flat, device=self.device)
        else:
            if flat.size and (np.nanmin(flat) < self.x[0] or np.nanmax(flat) > self.x[-1]):
                raise ValueError("A value in x_new is outside the interpolation range.")
            out = np.interp(flat, self.x, self.y)
        return out[0] if scalar else out.reshape(x_new.shape)
