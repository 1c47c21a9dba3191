"""Time one bootstrap replicate whose indices are NumPy's own MT19937 randint stream drawn on the device (lf_boot_bin_mt),
and check the first replicate against the host-drawn one.

    python tools/mt_time.py [n ...]        (default 1e6 1e7)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lumfuncmcmc_b200.engine import VeffEngine            # noqa: E402


def main():
    sizes = [int(float(a)) for a in sys.argv[1:]] or [10 ** 6, 10 ** 7]
    eng = VeffEngine(0)
    for n in sizes:
        rs = np.random.RandomState(3)
        lum = 40.0 + 4.0 * rs.random_sample(n)
        phi = rs.random_sample(n)
        edges = np.linspace(40.0, 44.0, 51)
        eng.bin_weights(lum, phi, edges)
        np.random.seed(11)
        state = np.random.get_state()
        idx = np.random.randint(n, size=n)
        host = eng.boot_bin(np.bincount(idx, minlength=n))
        eng.boot_mt_set_state(state)
        first = eng.boot_bin_mt()
        same = np.array_equal(first[0], host[0]) and np.array_equal(first[1], host[1])
        best, best_k = 1e9, 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            eng.boot_bin_mt()
            best = min(best, (time.perf_counter() - t0) * 1e3)
            best_k = min(best_k, eng.last_kernel_ms())
        print("mt19937 replicate at %d: %.3f ms wall, kernels %.3f ms, first replicate == host: %s" % (n, best, best_k, same),
              flush=True)
    eng.close()


if __name__ == '__main__':
    main()
