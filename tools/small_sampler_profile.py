import sys; sys.path.insert(0, ".")
import numpy as np
from lumfuncmcmc_b200 import synth
from lumfuncmcmc_b200.engine import LikelihoodEngine
from lumfuncmcmc_b200.sampler import DeviceEnsembleSampler
cat = synth.make_catalogue(10000, seed=4242)
inp = synth.direct_inputs(cat, nknots=4096, size_ln=101)
eng = LikelihoodEngine(inp, 'free')
th = synth.draw_thetas(inp, 'free', 100, seed=9, mode='near', scale=0.02)
s = DeviceEnsembleSampler(100, eng.ndim, eng, seed=17)
s.run_mcmc(th, 20)
print("device ms per step", s.device_ms / 20)
