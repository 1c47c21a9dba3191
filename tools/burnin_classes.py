"""How often does each walker class occur in a real run?  300 ensemble updates of 100 walkers from the reference's own
initialisation (reference lumfuncmcmc.py:436-446) on a 1e5-source catalogue; prints rejected / fast / literal walkers per
lnprob call and the kernel time.   python tools/burnin_classes.py"""
import sys, time; sys.path.insert(0, ".")
import numpy as np
from lumfuncmcmc_b200 import synth, configLF
from lumfuncmcmc_b200.lumfuncmcmc import LumFuncMCMC
from lumfuncmcmc_b200.sampler import EnsembleSampler
cat = synth.make_catalogue(100000, seed=2024, nfields=5)
m = LumFuncMCMC(cat['z'], flux=cat['flux'], flux_e=cat['flux_e'], Flim=list(cat['Flim']), alpha=cat['alpha'],
                Omega_0=list(cat['Omega_0']), Flim_lims=configLF.Flim_lims, alpha_lims=configLF.alpha_lims,
                sch_al=configLF.sch_al, Lstar=configLF.Lstar, phistar=configLF.phistar, fcmin=cat['fcmin'],
                min_comp_frac=0.0, field_names=cat['field_names'], field_ind=cat['field_ind'], nwalkers=100, nsteps=300)
np.random.seed(1)
pos = m.get_init_walker_values()
eng = m._engine('free')
log = []
def f(th):
    t0 = time.perf_counter(); out = m.lnprob(th); dt = time.perf_counter() - t0
    info = eng.last_call_info(); log.append((info['rejected'], info['fast'], info['literal'], dt * 1e3, eng.last_kernel_ms()))
    return out
s = EnsembleSampler(100, pos.shape[1], f, vectorize=True)
s.run_mcmc(pos, 300, rstate0=np.random.get_state())
L = np.array(log)
for a, b in ((0, 1), (1, 21), (21, 101), (101, 301), (301, 601)):
    x = L[a:b]
    print("calls %3d-%3d: rejected %.1f fast %.1f literal %.1f per call; kernel ms %.3f (max %.3f)" % (a, b, x[:,0].mean(), x[:,1].mean(), x[:,2].mean(), x[:,4].mean(), x[:,4].max()))
print("total kernel ms", L[:,4].sum(), "of which calls with literal walkers", L[L[:,2] > 0, 4].sum())
