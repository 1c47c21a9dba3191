"""Wall clock vs device time of the device-resident sampler at the BASELINE configs[0] size (1e4 sources x 100 walkers)
for several run lengths: separates the fixed cost of a run (allocation, capture, instantiation, downloads) from the
per-update cost.    python tools/small_sampler_profile.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lumfuncmcmc_b200 import synth                               # noqa: E402
from lumfuncmcmc_b200.engine import LikelihoodEngine             # noqa: E402

cat = synth.make_catalogue(10000, seed=4242)
inp = synth.direct_inputs(cat, nknots=4096, size_ln=101)
eng = LikelihoodEngine(inp, 'free')
th = synth.draw_thetas(inp, 'free', 100, seed=9, mode='near', scale=0.02)
eng.sampler_run(th, 20, 17)
for store in (True, False):
    for n in (0, 1, 100, 500, 2000, 8000):
        best, dev = 1e9, 0.0
        for _ in range(3):
            t0 = time.perf_counter()
            out = eng.sampler_run(th, n, 17, store_chain=store)
            dt = time.perf_counter() - t0
            if dt < best:
                best, dev = dt, out['device_ms']
        print("store_chain=%-5s nsteps=%5d  wall %8.2f ms  device (graph launches) %8.2f ms  -> %.4f ms wall per update" % (
            store, n, best * 1e3, dev, best * 1e3 / max(n, 1)))
