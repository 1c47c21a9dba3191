"""BASELINE.json configs[3] through the class: LumFuncMCMC(...).VeffLF() on N sources (weights, binned LF, nboot bootstrap
replicates with NumPy's own MT19937 stream drawn on the device), wall clock, for min_comp_frac = 0 and 0.5.
    python tools/veff_class_time.py [N] [nboot]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lumfuncmcmc_b200 import configLF, synth                  # noqa: E402
from lumfuncmcmc_b200.lumfuncmcmc import LumFuncMCMC           # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
nboot = int(sys.argv[2]) if len(sys.argv) > 2 else 100
cat = synth.make_catalogue(n, seed=3, nfields=5)
for mcf in (0.0, 0.5):
    t0 = time.perf_counter()
    m = LumFuncMCMC(cat['z'], flux=cat['flux'], flux_e=cat['flux_e'], Flim=list(cat['Flim']), alpha=cat['alpha'],
                    Omega_0=list(cat['Omega_0']), Flim_lims=configLF.Flim_lims, alpha_lims=configLF.alpha_lims,
                    sch_al=configLF.sch_al, Lstar=configLF.Lstar, phistar=configLF.phistar, fcmin=cat['fcmin'],
                    min_comp_frac=mcf, field_names=cat['field_names'], field_ind=cat['field_ind'], nbins=50, nboot=nboot)
    t_ctor = time.perf_counter() - t0
    np.random.seed(11)
    ts = []
    for rep in range(2):                     # second call: sample, u, rows and counts are already resident
        t0 = time.perf_counter()
        m.VeffLF()
        ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    phi = m.phifunc
    t_phi = time.perf_counter() - t0
    print("N=%d min_comp_frac=%.1f nboot=%d: constructor %.2f s; VeffLF() first call %.3f s, second call %.3f s; phifunc download %.3f s; "
          "%d sources with a volume, sum of bin counts %d" % (n, mcf, nboot, t_ctor, ts[0], ts[1], t_phi, int(np.count_nonzero(phi)), int(np.sum(m.bincounts))),
          flush=True)
    m.close()
