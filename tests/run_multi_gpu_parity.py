"""Run under torchrun (one rank per GPU): source-sharded lnprob over NCCL equals the single-GPU result and the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/run_multi_gpu_parity.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lumfuncmcmc_b200 import synth                          # noqa: E402
from lumfuncmcmc_b200.dist import ShardedLikelihood, WalkerShardedLikelihood, shard_inputs   # noqa: E402
from lumfuncmcmc_b200.engine import LikelihoodEngine        # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
ok = True
for kind in ('free', 'fixed', 'z'):
    cat = synth.make_catalogue(200000, seed=41, evolve=(0.3, -0.2) if kind == 'z' else None)
    inp = synth.direct_inputs(cat, nknots=2048, size_ln=101 if kind == 'free' else 201, tabulated=(kind != 'free'))
    th = np.concatenate([synth.draw_thetas(inp, kind, 96, seed=5), synth.draw_thetas(inp, kind, 32, seed=6, mode='prior')])
    like = ShardedLikelihood(shard_inputs(inp, rank, world), kind, device=local)
    got = like.lnprob(th)
    # the same exchange as one kernel over peer memory instead of NCCL: identical on every rank, equal to the NCCL sum
    p2p = ShardedLikelihood(shard_inputs(inp, rank, world), kind, device=local, exchange='p2p', wcap=256)
    for _ in range(3):                                     # repeated calls exercise both parities of the buffers
        got_p = p2p.lnprob(th)
    gathered = [None] * world
    dist.all_gather_object(gathered, got_p.tobytes())
    same_everywhere = all(g == gathered[0] for g in gathered)
    with np.errstate(invalid='ignore'):
        fin_p = np.isfinite(got)
        rel_p = np.max(np.abs(got_p[fin_p] - got[fin_p]) / np.abs(got[fin_p])) if fin_p.any() else 0.0
    p2p_ok = same_everywhere and np.array_equal(np.isneginf(got_p), np.isneginf(got)) and rel_p < 1e-13
    # device-resident sampler across the ranks (CUDA graph per update, peer-memory exchange inside the graph): every
    # rank ends with the same chain, equal to the host replay of the Philox stream driven by the sharded lnprob
    chain_ok = True
    if kind != 'fixed':
        from lumfuncmcmc_b200.sampler import philox_stretch_reference
        p0 = synth.draw_thetas(inp, kind, 64, seed=8, mode='near', scale=0.01)
        run = p2p.engine.sampler_run(p0, 12, seed=99)
        ref_chain, ref_lnp, _ = philox_stretch_reference(p2p.lnprob, p0, 12, 99)
        g2 = [None] * world
        dist.all_gather_object(g2, run['chain'].tobytes())
        chain_ok = all(g == g2[0] for g in g2) and np.array_equal(run['chain'], ref_chain) and np.array_equal(run['lnprob'], ref_lnp)
    p2p_ok = p2p_ok and chain_ok
    p2p.close()
    # walker sharding (every rank holds all sources): lnprob through NCCL all-gather and through the peer-memory exchange
    # equals one engine; the walker-sharded device-resident sampler gives the same chain on every rank = its host replay
    ws_n = WalkerShardedLikelihood(inp, kind, device=local)
    ws_p = WalkerShardedLikelihood(inp, kind, device=local, exchange='p2p', wcap=256)
    got_wn, got_wp = ws_n.lnprob(th), ws_p.lnprob(th)
    ws_ok = np.array_equal(got_wn, got_wp, equal_nan=True)
    ws_chain_ok = True
    if kind != 'fixed':
        from lumfuncmcmc_b200.sampler import philox_stretch_reference
        p0 = synth.draw_thetas(inp, kind, 64, seed=8, mode='near', scale=0.01)
        run = ws_p.sampler_run(p0, 12, seed=77)
        ref_chain, ref_lnp, _ = philox_stretch_reference(ws_p.lnprob, p0, 12, 77)
        g3 = [None] * world
        dist.all_gather_object(g3, run['chain'].tobytes())
        ws_chain_ok = all(g == g3[0] for g in g3) and np.array_equal(run['chain'], ref_chain) and np.array_equal(run['lnprob'], ref_lnp)
    ws_n.close()
    ws_p.close()
    if rank == 0:
        from oracle import lf_oracle
        single = LikelihoodEngine(inp, kind, device=local)
        one = single.lnprob(th)
        same_inf = np.array_equal(np.isneginf(got), np.isneginf(one))
        fin = np.isfinite(one)
        rel = np.max(np.abs(got[fin] - one[fin]) / np.abs(one[fin]))
        ref = lf_oracle.lnprob_batch(inp, kind, th[:12])
        f2 = np.isfinite(ref)
        rel_o = np.max(np.abs(got[:12][f2] - ref[f2]) / np.abs(ref[f2]))
        print("%s: world=%d  -inf sets equal=%s  max rel vs single GPU %.2e  vs oracle %.2e   peer-memory exchange: identical on "
              "all ranks=%s, max rel vs NCCL %.2e, multi-GPU device sampler chain == host replay on all ranks: %s"
              % (kind, world, same_inf, rel, rel_o, same_everywhere, rel_p, chain_ok))
        with np.errstate(invalid='ignore'):
            rel_w = np.max(np.abs(got_wp[fin] - one[fin]) / np.abs(one[fin]))
        ws_all = ws_ok and ws_chain_ok and np.array_equal(np.isneginf(got_wp), np.isneginf(one)) and rel_w < 1e-13
        print("   walker-sharded: NCCL all-gather == peer-memory gather: %s, max rel vs one engine %.2e, walker-sharded device sampler "
              "chain == host replay on all ranks: %s" % (ws_ok, rel_w, ws_chain_ok))
        ok = ok and same_inf and rel < 1e-12 and rel_o < 1e-10 and p2p_ok and ws_all
        single.close()
    like.close()
    dist.barrier()
if rank == 0:
    print("MULTI_GPU_PARITY", "OK" if ok else "FAIL")
dist.destroy_process_group()
