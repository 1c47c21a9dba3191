"""BASELINE.json configs[0] and configs[2] end to end through the unchanged command-line drivers on one GPU.

    python tests/run_config_end_to_end.py [0|2|all] [host|device]

configs[0]: run_lumfuncmcmc.py single-z Schechter + Fleming fit, synthetic 1e4-source catalogue drawn from known
            (logL*, logphi*, alpha) = (42.5, -2.0, -1.49), 100 walkers x 1000 steps.
configs[2]: run_lumfuncmcmc_z.py redshift-evolving LF, 1e6 sources with truth evolving in z, 512 walkers (a short run:
            the point is the wall clock per step and the set-up time, not a converged posterior).
Prints wall clocks (set-up / sampling / whole driver), the posterior medians against the truth, and for configs[0] the
time the reference algorithm (oracle port, 1 core, as the reference runs emcee) needs per lnprob call at the same size.
"""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lumfuncmcmc_b200 import synth                                  # noqa: E402
from lumfuncmcmc_b200.tableio import Table                          # noqa: E402


def write_catalogue(path, cat, line='OIII'):
    with open(path, 'w') as fh:
        fh.write('Field z ID %s_flux %s_flux_e\n' % (line, line))
        i = 0
        for k, name in enumerate(cat['field_names']):
            zz, ff, fe = cat['z'][k], cat['flux'][k], cat['flux_e'][k]
            ids = np.arange(i, i + len(zz))
            np.savetxt(fh, np.column_stack([zz, ids, ff, fe]), fmt=name + ' %.17g %d %.17g %.17g')
            i += len(zz)


def run_config0(backend):
    import run_lumfuncmcmc
    work = tempfile.mkdtemp(prefix='lfcfg0_')
    cwd = os.getcwd()
    os.chdir(work)
    try:
        cat = synth.make_catalogue(10000, seed=2024, nfields=5)
        write_catalogue('cat.dat', cat)
        os.environ['LF_SAMPLER'] = backend
        np.random.seed(7)
        t0 = time.perf_counter()
        m = run_lumfuncmcmc.main(['-f', 'cat.dat', '-o', 'cfg0.dat', '-nw', '100', '-ns', '1000', '-nboot', '100'])
        wall = time.perf_counter() - t0
        names = m.get_param_names()
        med = np.median(m.samples[:, :-1], axis=0)
        out = {"config": "BASELINE.json configs[0]: 1e4 sources, 100 walkers x 1000 steps, driver end to end", "sampler": backend,
               "driver_wall_s": wall, "acceptance": float(np.mean(m.sampler.acceptance_fraction)),
               "posterior_medians": {n: float(v) for n, v in zip(names, med)},
               "truth": {"logL*": 42.5, "logphi*": -2.0, "alpha": -1.49, "Flim": [float(x) for x in cat['Flim']], "alpha_c": float(cat['alpha'])},
               "outputs": sorted(os.listdir('LFMCMCOut'))}
        # the reference algorithm at the same size, one call at a time on one core (how the reference runs emcee)
        from oracle import lf_oracle
        model = lf_oracle.make_model(m.engine_inputs(), 'free')
        th = m.samples[:20, :-1]
        model.lnprob(th[0])
        t0 = time.perf_counter()
        ref = np.array([model.lnprob(t) for t in th])
        per_call = (time.perf_counter() - t0) / len(th)
        got = m.lnprob(th)
        out["reference_s_per_lnprob_call_1core"] = per_call
        out["reference_projected_run_s"] = per_call * 100 * 1001
        out["max_rel_diff_vs_reference_on_posterior_samples"] = float(np.max(np.abs(got - ref) / np.abs(ref)))
        m.close()
        return out
    finally:
        os.chdir(cwd)
        shutil.rmtree(work, ignore_errors=True)


def run_config2(backend, n=1000000, walkers=512, steps=20):
    import run_lumfuncmcmc_z
    work = tempfile.mkdtemp(prefix='lfcfg2_')
    cwd = os.getcwd()
    os.chdir(work)
    try:
        cat = synth.make_catalogue(n, seed=2025, nfields=5, evolve=(0.3, -0.2))
        t0 = time.perf_counter()
        write_catalogue('catz.dat', cat)
        t_write = time.perf_counter() - t0
        os.environ['LF_SAMPLER'] = backend
        np.random.seed(8)
        t0 = time.perf_counter()
        m = run_lumfuncmcmc_z.main(['-f', 'catz.dat', '-o', 'cfg2.dat', '-nw', str(walkers), '-ns', str(steps), '-nboot', '10'])
        wall = time.perf_counter() - t0
        out = {"config": "BASELINE.json configs[2]: z-evolving LF, %g sources, %d walkers x %d steps, driver end to end" % (n, walkers, steps),
               "sampler": backend, "driver_wall_s": wall, "catalogue_write_s": t_write,
               "acceptance": float(np.mean(m.sampler.acceptance_fraction)),
               "sampler_device_ms_per_step": getattr(m.sampler, 'device_ms', 0.0) / steps,
               "outputs": sorted(os.listdir('LFMCMCzOut'))}
        m.close()
        return out
    finally:
        os.chdir(cwd)
        shutil.rmtree(work, ignore_errors=True)


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    backend = sys.argv[2] if len(sys.argv) > 2 else 'host'
    if which in ('0', 'all'):
        print(json.dumps(run_config0(backend)))
    if which in ('2', 'all'):
        print(json.dumps(run_config2(backend)))
