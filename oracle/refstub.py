"""TEST INFRASTRUCTURE ONLY -- stub-import of the UNMODIFIED reference (authoring container only).

`/root/reference` holds the reference's six Python files; they import packages that are not in this
image (emcee, astropy, uncertainties, matplotlib, corner, seaborn, lmfit) and the removed name
``scipy.integrate.trapz``.  This module registers minimal stand-ins in ``sys.modules`` and then imports
the reference's own files, so that golden vectors can be produced by the reference's own bytecode
(SURVEY.md Appendix C).  Only ``astropy.cosmology.LambdaCDM`` and ``uncertainties.unumpy`` need working
behaviour; the former feeds set-up tables only (never the per-step arithmetic).

`/root/reference` does not exist on the GPU box: nothing under ``tests/`` marked ``gpu``, ``smoke()`` or
``bench.py`` may import this module.  It is used by ``oracle/make_golden.py`` and by the CPU-side tests that
re-validate ``oracle/lf_oracle.py`` when the reference happens to be present.
"""
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get('LF_REFERENCE_DIR', '/root/reference')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, 'lumfuncmcmc.py'))


class _UArr:
    """First-order error propagation on arrays: just enough of ``uncertainties.unumpy.uarray``."""
    __array_ufunc__ = None

    def __init__(self, n, s):
        self.n = np.asarray(n, dtype=np.float64)
        self.s = np.asarray(s, dtype=np.float64)

    def __mul__(self, other):
        other = np.asarray(other, dtype=np.float64)
        return _UArr(other * self.n, np.abs(other) * self.s)

    __rmul__ = __mul__

    def __truediv__(self, other):
        other = np.asarray(other, dtype=np.float64)
        return _UArr(self.n / other, self.s / np.abs(other))

    def __rpow__(self, base):
        val = base ** self.n
        return _UArr(val, np.abs(val * np.log(base)) * self.s)


def _install_stubs():
    import scipy.integrate
    if not hasattr(scipy.integrate, 'trapz'):
        scipy.integrate.trapz = scipy.integrate.trapezoid

    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    if root not in sys.path:
        sys.path.insert(0, root)
    from lumfuncmcmc_b200 import cosmology as _cosmology

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, name):
            return _Anything()

        def __iter__(self):
            return iter(())

    class _Unit:
        """supports ``70 * u.km / u.s / u.Mpc`` and ``2.725 * u.K`` -> objects carrying ``.value``"""

        def __init__(self, value=1.0):
            self.value = value

        def __rmul__(self, other):
            return _Unit(float(other))

        def __truediv__(self, other):
            return _Unit(self.value)

    if 'emcee' not in sys.modules:
        mod('emcee', EnsembleSampler=_Anything)
    unumpy = mod('uncertainties.unumpy',
                 uarray=lambda n, s: _UArr(n, s),
                 log10=lambda u: _UArr(np.log10(u.n), np.abs(u.s / (u.n * np.log(10.0)))),
                 exp=lambda u: _UArr(np.exp(u.n), np.abs(np.exp(u.n)) * u.s),
                 nominal_values=lambda u: u.n,
                 std_devs=lambda u: u.s)
    mod('uncertainties', unumpy=unumpy, ufloat=_Anything)
    plt = mod('matplotlib.pyplot')
    plt.__getattr__ = lambda name: _Anything()
    mod('matplotlib', use=lambda *a, **k: None, rcParams={}, pyplot=plt)
    mod('corner', corner=_Anything())
    sns = mod('seaborn')
    sns.set_context = sns.set_style = sns.set_palette = lambda *a, **k: None
    sns.xkcd_palette = lambda *a, **k: []
    sns.color_palette = lambda *a, **k: [(0.0, 0.0, 0.0)]
    mod('lmfit', Model=_Anything)
    mod('mpl_toolkits')
    mod('mpl_toolkits.axes_grid1', make_axes_locatable=_Anything())

    def _lcdm(H0=70.0, Tcmb0=2.725, Om0=0.3, Ode0=0.7, **kw):
        return _cosmology.LambdaCDM(H0=H0, Tcmb0=Tcmb0, Om0=Om0, Ode0=Ode0).as_astropy_like()

    astropy = mod('astropy')
    astropy.table = mod('astropy.table', Table=_Anything)
    astropy.units = mod('astropy.units', km=_Unit(), s=_Unit(), Mpc=_Unit(), K=_Unit())
    astropy.cosmology = mod('astropy.cosmology', LambdaCDM=_lcdm)


_loaded = {}


def load_reference():
    """Return (VmaxLumFunc, lumfuncmcmc, lumfuncmcmc_z) imported unmodified from the reference tree."""
    if _loaded:
        return _loaded['V'], _loaded['lf'], _loaded['lfz']
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_DIR)
    _install_stubs()
    sys.dont_write_bytecode = True          # the reference directory is read-only
    if REFERENCE_DIR not in sys.path:
        sys.path.append(REFERENCE_DIR)
    import importlib
    # import under private names so the product's same-named modules are never shadowed
    saved = {k: sys.modules.pop(k) for k in ('VmaxLumFunc', 'lumfuncmcmc', 'lumfuncmcmc_z', 'configLF')
             if k in sys.modules}
    try:
        V = importlib.import_module('VmaxLumFunc')
        lf = importlib.import_module('lumfuncmcmc')
        lfz = importlib.import_module('lumfuncmcmc_z')
    finally:
        for k in ('VmaxLumFunc', 'lumfuncmcmc', 'lumfuncmcmc_z', 'configLF'):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
        if REFERENCE_DIR in sys.path:
            sys.path.remove(REFERENCE_DIR)
    _loaded.update(V=V, lf=lf, lfz=lfz)
    return V, lf, lfz
