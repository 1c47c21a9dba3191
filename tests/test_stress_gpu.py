"""GPU: robustness -- odd ensemble sizes, repeated context creation (no device-memory leak), long graph-replayed runs."""
import numpy as np
import pytest

from lumfuncmcmc_b200 import synth
from oracle import lf_oracle

pytestmark = pytest.mark.gpu


def _inputs(n=1500, seed=3, kind='free'):
    cat = synth.make_catalogue(n, seed=seed, nfields=3, evolve=(0.3, -0.2) if kind == 'z' else None)
    return synth.direct_inputs(cat, nknots=256, size_ln=41 if kind == 'free' else 61, tabulated=(kind != 'free'))


@pytest.mark.parametrize('W', [1, 2, 31, 33, 100, 4100])
def test_odd_ensemble_sizes(W):
    from lumfuncmcmc_b200.engine import LikelihoodEngine
    inp = _inputs()
    th = np.concatenate([synth.draw_thetas(inp, 'free', max(W - W // 4, 1), seed=1, mode='near', scale=0.02),
                         synth.draw_thetas(inp, 'free', W // 4, seed=2, mode='prior')])[:W]
    eng = LikelihoodEngine(inp, 'free', device=0)
    got = eng.lnprob(th)
    idx = np.unique(np.linspace(0, W - 1, min(W, 24)).astype(int))
    ref = lf_oracle.lnprob_batch(inp, 'free', th[idx])
    assert np.array_equal(np.isneginf(got[idx]), np.isneginf(ref))
    fin = np.isfinite(ref)
    assert np.max(np.abs(got[idx][fin] - ref[fin]) / np.abs(ref[fin])) < 1e-10
    # growing and shrinking the batch on the same context re-plans the work items
    again = eng.lnprob(th[:max(1, W // 2)])
    assert np.array_equal(again, got[:max(1, W // 2)], equal_nan=True)
    eng.close()


def test_repeated_contexts_do_not_leak_device_memory():
    import torch
    from lumfuncmcmc_b200.engine import LikelihoodEngine, VeffEngine
    inp = _inputs(20000)
    th = synth.draw_thetas(inp, 'free', 64, seed=1, mode='near', scale=0.02)

    def cycle():
        e = LikelihoodEngine(inp, 'free', device=0, compress=True)
        e.lnprob(th)
        e.sampler_run(th, 3, seed=1)
        e.close()
        v = VeffEngine()
        lum = np.linspace(41.0, 43.0, 5000)
        v.bin_weights(lum, np.ones(5000), np.linspace(41.0, 43.0, 11))
        v.boot_bin_device(1, 0)
        v.close()
    cycle()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(25):
        cycle()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 8 << 20, "device memory shrank by %.1f MiB over 25 create/destroy cycles" % ((free0 - free1) / 2 ** 20)


def test_long_graph_replayed_run_stays_consistent():
    """5000 captured updates: the chain's stored log-posteriors are the engine's values at the stored positions at the
    start, in the middle and at the end, and the acceptance fraction is that of a healthy stretch-move run."""
    from lumfuncmcmc_b200.engine import LikelihoodEngine
    inp = _inputs(3000)
    eng = LikelihoodEngine(inp, 'free', device=0)
    p0 = synth.draw_thetas(inp, 'free', 40, seed=4, mode='near', scale=0.01)
    run = eng.sampler_run(p0, 5000, seed=77)
    for t in (0, 2500, 4999):
        want = eng.lnprob(run['chain'][t])
        assert np.allclose(run['lnprob'][t], want, rtol=1e-12, atol=0, equal_nan=True)
    acc = run['naccepted'] / 5000.0
    assert 0.15 < acc.mean() < 0.7 and np.all(np.isfinite(run['lnprob'][-1]))
    eng.close()
