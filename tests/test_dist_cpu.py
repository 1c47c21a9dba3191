"""CPU, gloo, world_size 2: host-side logic of source sharding -- shard bookkeeping, quadrature-share assignment and
the SUM all-reduce semantics (-inf propagates) -- with the NumPy oracle standing in for the per-rank GPU kernels."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from lumfuncmcmc_b200 import synth
    from lumfuncmcmc_b200.dist import reduce_partials, shard_inputs
    from oracle import lf_oracle
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    cat = synth.make_catalogue(3000, seed=2, nfields=3)
    inp = synth.direct_inputs(cat, nknots=256, size_ln=41)
    th = np.concatenate([synth.draw_thetas(inp, 'free', 6, seed=1), synth.draw_thetas(inp, 'free', 6, seed=2, mode='prior')])
    mine = shard_inputs(inp, rank, world)
    assert mine['field_ind'][-1] == len(mine['lum'])
    # per-rank partial = shard's source sum (zero-weight quadrature) - quadrature of "my" walkers only
    zero_quad = dict(mine, volume_part=np.zeros_like(mine['volume_part']))
    part = lf_oracle.lnprob_batch(zero_quad, 'free', th)
    nosrc = dict(inp, lum=inp['lum'][:0], z=inp['z'][:0], field_ind=np.zeros_like(inp['field_ind']))
    quad = lf_oracle.lnprob_batch(nosrc, 'free', th)                 # = -fullint (or -inf outside the prior)
    own = (np.arange(len(th)) % world) == rank
    with np.errstate(invalid='ignore'):
        part = np.where(own, part + quad, part)
    t = torch.from_numpy(part.copy())
    reduce_partials(t)
    full = lf_oracle.lnprob_batch(inp, 'free', th)
    ret[rank] = (t.numpy().copy(), full, len(mine['lum']))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_source_sharding_sums_to_the_full_lnprob():
    world, port = 2, 29000 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    (a, full, n0), (b, _, n1) = ret[0], ret[1]
    assert n0 + n1 == 3000
    assert np.array_equal(a, b, equal_nan=True)                      # every rank holds the reduced vector
    assert np.array_equal(np.isneginf(a), np.isneginf(full))
    fin = np.isfinite(full)
    assert fin.sum() >= 6
    assert np.max(np.abs(a[fin] - full[fin]) / np.abs(full[fin])) < 1e-12


def test_shard_bounds_cover_everything_once():
    from lumfuncmcmc_b200.dist import shard_bounds
    for n in (0, 1, 7, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def _gather_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from lumfuncmcmc_b200.dist import gather_walker_results, walker_slice
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    out = {}
    for W in (1, 2, 7, 100):
        lo, hi = walker_slice(W, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float64) * 1.5 - 3.0
        if hi > lo and lo == 0:
            local[0] = float('-inf')
        out[W] = gather_walker_results(local, W).numpy().copy()
    ret[rank] = out
    dist.barrier()
    dist.destroy_process_group()


def test_walker_sharding_gathers_the_full_vector_on_every_rank():
    world, port = 2, 31000 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gather_worker, args=(world, port, ret), nprocs=world, join=True)
    for W in (1, 2, 7, 100):
        want = np.arange(W) * 1.5 - 3.0
        want[0] = -np.inf
        for r in range(world):
            assert np.array_equal(ret[r][W], want)


def test_peer_geometry_mismatch_is_refused():
    """Every rank's peer buffer must be created with the same (world, wcap): the slot offsets are computed from them."""
    from lumfuncmcmc_b200.dist import check_peer_geometry
    ok = [(b'h0', 0, 2, 4096), (b'h1', 1, 2, 4096)]
    check_peer_geometry(ok, 2, 4096)
    for bad in ([(b'h0', 0, 2, 4096), (b'h1', 1, 2, 2048)],        # other capacity
                [(b'h0', 0, 2, 4096), (b'h1', 1, 4, 4096)],        # other world
                [(b'h0', 0, 2, 4096), (b'h1', 0, 2, 4096)]):       # duplicate rank
        with pytest.raises(RuntimeError, match='disagree'):
            check_peer_geometry(bad, 2, 4096)
