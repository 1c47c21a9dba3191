"""Resident 1/V_eff weights pass (k_veff_res) and bootstrap replicate over a range of sample sizes: event-timed kernel ms per
size (fixed cost vs marginal rate).  Run under `ncu --metrics gpu__time_duration.sum -k regex:k_veff` for kernel-only times.
    python tools/veff_sweep.py [nbins]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lumfuncmcmc_b200.engine import VeffEngine   # noqa: E402

nbins = int(sys.argv[1]) if len(sys.argv) > 1 else 50
rng = np.random.default_rng(4)
ve = VeffEngine()
for n in (1000000, 2000000, 5000000, 10000000, 20000000, 40000000):
    lum = rng.uniform(40.9, 44.0, n)
    flux = 10 ** rng.uniform(-17.2, -14.5, n)
    fi = np.array([0, n // 5, 2 * n // 5, 3 * n // 5, 4 * n // 5, n], dtype=np.int64)
    edges = np.linspace(lum.min() * 1.001, lum.max(), nbins + 1)
    ve.veff_set_sample(flux, lum, fi)
    ve.veff_bin_resident([2.72, 3.61, 2.55, 3.31, 3.30], 4.56, 0.1, 1.9e6, 3.0e10, edges)
    best = min(ve.veff_bin_resident([2.72, 3.61, 2.55, 3.31, 3.30], 4.56, 0.1, 1.9e6, 3.0e10, edges) and ve.last_kernel_ms() for _ in range(4))
    mult = np.bincount(rng.integers(0, n, n), minlength=n)
    bb = min(ve.boot_bin(mult) and ve.last_kernel_ms() for _ in range(4))
    print("N=%9d  weights %.4f ms (%.0f GB/s of 26 B)   replicate %.4f ms (%.0f GB/s of 14 B)" % (
        n, best, 26.0 * n / (best * 1e-3) / 1e9, bb, 14.0 * n / (bb * 1e-3) / 1e9), flush=True)
