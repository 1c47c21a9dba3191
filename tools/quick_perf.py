"""Quick device-side throughput probe (not the bench): python tools/quick_perf.py [N] [W] [kind]"""
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lumfuncmcmc_b200 import synth
from lumfuncmcmc_b200.engine import LikelihoodEngine

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
kind = sys.argv[3] if len(sys.argv) > 3 else 'free'
prec = sys.argv[4] if len(sys.argv) > 4 else 'f64'
t0 = time.time()
cat = synth.make_catalogue(N, seed=1, evolve=(0.3, -0.2) if kind == 'z' else None)
inp = synth.direct_inputs(cat, nknots=4096, size_ln=101 if kind == 'free' else 201, tabulated=(kind != 'free'))
print("catalogue+inputs %.1fs" % (time.time() - t0))
t0 = time.time()
eng = LikelihoodEngine(inp, kind, precision=prec)
print("engine set-up %.2fs" % (time.time() - t0))
rate, ms = eng.fp64_peak(20000)
print("fp64 peak: %.3e DFMA/s (%.2f ms) -> %.1f per clk per SM at 1.965 GHz x 148" % (rate, ms, rate / 1.965e9 / 148))
th = synth.draw_thetas(inp, kind, W, seed=3, mode='near', scale=0.02)
for it in range(5):
    t0 = time.time()
    out = eng.lnprob(th)
    wall = time.time() - t0
    kms = eng.last_kernel_ms()
    print("call %d: wall %.2f ms kernels %.2f ms  terms/s (kernel) %.3e  (wall) %.3e  info %s" % (
        it, wall * 1e3, kms, N * W / (kms * 1e-3), N * W / wall, eng.last_call_info()))
print(out[:4])
rate, ms = eng.mufu_peak(20000)
print("mufu peak: %.3e ex2/s (%.2f ms) -> %.1f per clk per SM at 1.965 GHz x 148" % (rate, ms, rate / 1.965e9 / 148))
rate, ms = eng.fp64_peak(200000)
print("fp64 peak (long): %.3e DFMA/s (%.2f ms)" % (rate, ms))
