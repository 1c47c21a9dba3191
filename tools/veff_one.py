"""A few resident 1/V_eff weight passes at one size (for ncu captures of k_veff_res).   python tools/veff_one.py [N] [nbins]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lumfuncmcmc_b200.engine import VeffEngine   # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
nbins = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rng = np.random.default_rng(4)
lum = rng.uniform(40.9, 44.0, n)
flux = 10 ** rng.uniform(-17.2, -14.5, n)
fi = np.array([0, n // 5, 2 * n // 5, 3 * n // 5, 4 * n // 5, n], dtype=np.int64)
edges = np.linspace(lum.min() * 1.001, lum.max(), nbins + 1)
ve = VeffEngine()
ve.veff_set_sample(flux, lum, fi)
for _ in range(4):
    ve.veff_bin_resident([2.72, 3.61, 2.55, 3.31, 3.30], 4.56, 0.1, 1.9e6, 3.0e10, edges)
    print("N=%d kernel %.4f ms" % (n, ve.last_kernel_ms()))
