"""BASELINE.json configs 2-4 on one GPU: device throughput table (not the bench line).
    python tools/config_sweep.py [quick]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lumfuncmcmc_b200 import synth                           # noqa: E402
from lumfuncmcmc_b200.engine import LikelihoodEngine, VeffEngine   # noqa: E402

quick = len(sys.argv) > 1


# FP64-pipe instructions per (walker, source) term and per (walker, quadrature point) of the fast kernels (SASS counts,
# tools/sass_loop_mix.py): the quadrature is real work that "terms/s" does not count -- at 1e5 sources it is the larger half
FP64_TERM = {'free': 22, 'z': 9, 'fixed': 0}
FP64_POINT = {'free': 33, 'z': 11, 'fixed': 11}
PEAK = {}


def run(kind, n, walkers, prec='f64'):
    cat = synth.make_catalogue(n, seed=1, evolve=(0.3, -0.2) if kind == 'z' else None)
    inp = synth.direct_inputs(cat, nknots=4096, size_ln=101 if kind == 'free' else 201, tabulated=(kind != 'free'))
    eng = LikelihoodEngine(inp, kind, precision=prec)
    if 'dfma' not in PEAK:
        PEAK['dfma'] = max(eng.fp64_peak(100000)[0] for _ in range(3))
    S = len(inp['zarr'])
    npoints = (len(inp['Flim']) if kind == 'free' else 1) * S * S          # FIXED / Z: identical field grids are merged
    for W in walkers:
        th = synth.draw_thetas(inp, kind, W, seed=3, mode='near', scale=0.02)
        for _ in range(3):
            eng.lnprob(th)
        ts, ws = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            eng.lnprob(th)
            ws.append(time.perf_counter() - t0)
            ts.append(eng.last_kernel_ms() * 1e-3)
        k, w = min(ts), min(ws)
        frac_terms = n * W * FP64_TERM[kind] / k / PEAK['dfma']
        frac_all = (n * W * FP64_TERM[kind] + npoints * W * FP64_POINT[kind]) / k / PEAK['dfma']
        fr = "%.2f / %.2f" % (frac_terms, frac_all) if prec == 'f64' else "-"
        print("| %-5s | %s | %8.0e | %5d | %9.3f | %9.3f | %10.3e | %10.3e | %s |" % (kind, prec, n, W, k * 1e3, w * 1e3, n * W / k, n * W / w, fr),
              flush=True)
    eng.close()


print("| model | loop arithmetic | sources | walkers | kernel ms | host-call ms | terms/s (kernels) | terms/s (host API) | fraction of the measured DFMA rate: terms only / terms + quadrature |")
print("|---|---|---|---|---|---|---|---|---|")
sizes = [100000, 1000000] if quick else [100000, 1000000, 10000000]
for n in sizes:
    run('free', n, [64, 256, 1024, 4096])
    run('free', n, [64, 256, 1024, 4096], 'f32')
run('z', 1000000, [512])
run('z', 1000000, [512], 'f32')
run('fixed', 1000000, [512])
run('free', 10000, [100])

# config 4: 1/V_eff binned LF
n = 1000000 if quick else 10000000
rng = np.random.default_rng(4)
lum = rng.uniform(40.9, 44.0, n)
flux = 10 ** rng.uniform(-17.2, -14.5, n)
fi = np.array([0, n // 5, 2 * n // 5, 3 * n // 5, 4 * n // 5, n], dtype=np.int64)
edges = np.linspace(lum.min() * 1.001, lum.max(), 51)
ve = VeffEngine()
flim = [2.72, 3.61, 2.55, 3.31, 3.30]
for _ in range(2):
    t0 = time.perf_counter()
    phi, counts, sums = ve.veff_bin(flux, lum, fi, flim, 4.56, 0.1, 1.9e6, 3.0e10, edges)
    wall = time.perf_counter() - t0
kms = ve.last_kernel_ms()
want = np.histogram(lum[(lum >= edges[0]) & (lum < edges[-1])], bins=edges)[0]
print("\nVeff, host-buffer call (lf_veff_bin: uploads flux and lum, downloads the weights -- PCIe-bound): N=%d nbins=50  kernel %.3f ms  "
      "host call %.1f ms  counts bit-exact: %s" % (n, kms, wall * 1e3, np.array_equal(counts, want)))
# the route VeffLF takes: sample resident on the device, weights stay there
t0 = time.perf_counter()
ve.veff_set_sample(flux, lum, fi)
up = time.perf_counter() - t0
best_k, best_w = 1e9, 1e9
for _ in range(5):
    t0 = time.perf_counter()
    _, counts_r, sums_r = ve.veff_bin_resident(flim, 4.56, 0.1, 1.9e6, 3.0e10, edges)
    best_w = min(best_w, time.perf_counter() - t0)
    best_k = min(best_k, ve.last_kernel_ms())
print("Veff, resident sample (lf_veff_set_sample once: %.1f ms; then lf_veff_bin_resident): kernels %.4f ms (%.3e sources/s, %.1f GB/s "
      "of 26 B/source: flux, u, row in; phi out)  host call %.3f ms  counts bit-exact: %s  sums equal to the host-buffer call: %s"
      % (up * 1e3, best_k, n / (best_k * 1e-3), 26.0 * n / (best_k * 1e-3) / 1e9, best_w * 1e3, np.array_equal(counts_r, want),
         np.allclose(sums_r, sums, rtol=1e-13, atol=0)))
mult = np.bincount(rng.integers(0, n, n), minlength=n)
t0 = time.perf_counter()
bc, bs = ve.boot_bin(mult)
wall = time.perf_counter() - t0
kms = ve.last_kernel_ms()
print("bootstrap replicate, host-drawn multiplicities (lf_boot_bin: uploads 4 B per source): kernel %.3f ms (%.1f GB/s of 14 B/source: "
      "row i16 + phi f64 + multiplicity i32)  host call %.1f ms  counts sum %d" % (kms, 14.0 * n / (kms * 1e-3) / 1e9, wall * 1e3, bc.sum()))
