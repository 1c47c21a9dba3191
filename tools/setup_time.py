"""Set-up (constructor + engine creation) wall clock at N sources, device path vs host path.   python tools/setup_time.py [N]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lumfuncmcmc_b200 import configLF, setup_gpu, synth        # noqa: E402
import lumfuncmcmc_b200.lfbase as lfbase                       # noqa: E402
from lumfuncmcmc_b200.lumfuncmcmc import LumFuncMCMC           # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
cat = synth.make_catalogue(n, seed=3, nfields=5)


PHASES = ['defineFlimOmArr', 'getRoot', 'setDLdVdz', '_fluxes_and_luminosities', 'setOmegaLz', 'setlnsimple']
phase_s = {}


def _timed(cls, name):
    fn = getattr(cls, name)

    def wrap(self, *a, **k):
        t0 = time.perf_counter()
        out = fn(self, *a, **k)
        phase_s[name] = phase_s.get(name, 0.0) + time.perf_counter() - t0
        return out
    setattr(cls, name, wrap)


for _name in PHASES:
    _timed(LumFuncMCMC, _name)


def build():
    phase_s.clear()
    t0 = time.perf_counter()
    m = LumFuncMCMC(cat['z'], flux=cat['flux'], flux_e=cat['flux_e'], Flim=list(cat['Flim']), alpha=cat['alpha'],
                    Omega_0=list(cat['Omega_0']), Flim_lims=configLF.Flim_lims, alpha_lims=configLF.alpha_lims,
                    sch_al=configLF.sch_al, Lstar=configLF.Lstar, phistar=configLF.phistar, fcmin=cat['fcmin'],
                    min_comp_frac=0.0, field_names=cat['field_names'], field_ind=cat['field_ind'])
    t1 = time.perf_counter()
    th = np.array([42.5, -2.0, -1.49] + list(cat['Flim']) + [cat['alpha']])
    v = m.lnprob(th)                                           # creates the engine (uploads, derived arrays, statistics)
    t2 = time.perf_counter()
    m.close()
    return t1 - t0, t2 - t1, v


c, e, v = build()
c, e, v = build()                                              # second build: module loads and first-touch costs are paid
print("N=%d  device set-up path: constructor %.2f s, engine creation + first lnprob %.2f s, lnprob %.6f" % (n, c, e, v))
print("   phases: " + ", ".join("%s %.3f" % (k, phase_s[k]) for k in PHASES if k in phase_s))
lfbase.gpu_count = lambda: 0
setup_gpu.gpu_count = lambda: 0
c, e, v2 = build()
print("   phases: " + ", ".join("%s %.3f" % (k, phase_s[k]) for k in PHASES if k in phase_s))
print("N=%d  host set-up path:   constructor %.2f s, engine creation + first lnprob %.2f s, lnprob %.6f  (rel diff %.1e)"
      % (n, c, e, v2, abs(v - v2) / abs(v2)))
