#!/usr/bin/env python
"""bench.py -- walker x source lnL terms/s of the batched lnprob on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference algorithm on the host cores)

Headline workload = the north_star target sentence on BASELINE.json configs[1]'s corner: free-completeness single-z model
(ndim 9, K = 5 fields, S = 101), a 10^7-source synthetic catalogue x 1024 walkers, FP64.  A "step" is one batched lnprob
call over the whole ensemble.  At N GPUs the SAME 10^7 sources are sharded over the ranks (strong scaling, 10^7 / N per
GPU), walkers replicated, the quadrature split by walker, one all-reduce of W doubles per step.  `--scaling weak` keeps
10^7 sources per GPU instead.  Every line also carries, as sub-results measured in the same run:
  weak            10^7 sources PER GPU x 1024 walkers (N > 1; at N = 1 it is the headline itself)
  config4         BASELINE.json configs[4]: 10^8 sources in total x 2048 walkers, source-sharded (N > 1, as the config names)
  walker_sharded  10^5 sources x 4096 walkers, every rank holds all sources and evaluates W / N walkers (small-N regime)
  parity          what was checked in THIS run (see check_parity): all-reduce vs rank-ordered host sum of the un-reduced
                  partials, full-size oracle check of 2 walkers (every rank runs the oracle on its own shard), oracle check
                  of >= 8 walkers on a 10^6-source sub-shard, per-bin 1/V_eff counts vs NumPy at the full shard size

value  : terms/s, inputs resident in HBM (theta on the device), CUDA events on the launching stream, max over ranks.
e2e    : same metric through the public host API (ShardedLikelihood.lnprob: pinned-host theta -> H2D -> kernels
         -> all-reduce -> D2H of W doubles -> sync), host wall clock, max over ranks.
roofline: the loop is FP64-FMA-pipe bound (no tensor cores, HBM traffic ~0.02 B/term): achieved = terms/s/GPU x 22
         FP64-pipe instructions per term (counted in the SASS of k_main<false>) x 2 FLOP, against the register-only
         DFMA rate measured live on the same GPU (lf_fp64_peak, best of 3) x 2 FLOP.  HBM figures are reported beside it.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WARMUP_MIN_S = 0.3      # minimum duration of the untimed warm-up before a device-timed region
FP64_INSTR_PER_TERM_BY_KIND = {'free': 22, 'z': 9, 'fixed': 0}    # DFMA/DADD/DMUL per (walker, source) term in k_main<false, MODEL>
                                                                    # (tools/sass_loop_mix.py); fixed: sufficient statistics only
MUFU_PER_TERM_BY_KIND = {'free': 4, 'z': 1, 'fixed': 0}           # FP32 mode: rsqrt, lg2, ex2, rcp / one ex2
BYTES_PER_SOURCE = 16             # (log10 flux, flux) or (lum, z) per source per sweep
KIND_NAMES = {'free': 'free-completeness single-z', 'fixed': 'fixed-completeness single-z', 'z': 'redshift-evolving'}
ZRANGE = (1.16, 1.90)             # redshift range of the synthetic catalogue = range of the quadrature grid on every rank


def metric_name(kind, precision='f64'):
    return "walker x source lnL terms/sec (batched lnprob, %s model, %s)" % (KIND_NAMES[kind], "FP64" if precision == 'f64' else "FP32 loop")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='engine', choices=['engine', 'reference'])
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'],
                    help="strong: --nsources is the catalogue size, sharded over the GPUs; weak: --nsources per GPU")
    ap.add_argument('--nsources', type=float, default=1.0e7, help='sources in total (strong) or per GPU (weak)')
    ap.add_argument('--walkers', type=int, default=1024)
    ap.add_argument('--kind', default='free', choices=['free', 'fixed', 'z'])
    ap.add_argument('--precision', default='f64', choices=['f64', 'f32'], help='arithmetic of the walker x source loop')
    ap.add_argument('--exchange', default='nccl', choices=['nccl', 'p2p'],
                    help="multi-GPU sum of the per-walker partials: NCCL all-reduce or the engine's peer-memory kernel")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='headline only (no sub-results, samplers, V_eff): ncu / quick runs')
    ap.add_argument('--prior-draws', action='store_true', help='walkers ~ U(prior) instead of a converged ensemble')
    return ap.parse_args()


def build_inputs(n, kind, seed, lum_floor=None):
    from lumfuncmcmc_b200 import synth
    cat = synth.make_catalogue(n, seed=seed, evolve=(0.3, -0.2) if kind == 'z' else None, zmin=ZRANGE[0], zmax=ZRANGE[1])
    return synth.direct_inputs(cat, nknots=4096, size_ln=101 if kind == 'free' else 201, tabulated=(kind != 'free'),
                               zrange=ZRANGE, lum_floor=lum_floor)


def sample_inputs(inp, n_sample):
    """Bounded sample of the same workload for the CPU legs: the first n_sample/N of every field."""
    from lumfuncmcmc_b200.dist import shard_inputs
    n = len(inp['lum'])
    if n_sample >= n:
        return inp
    world = max(1, int(round(n / n_sample)))
    return shard_inputs(inp, 0, world)


class ClockSampler:
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.path = gpu, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """Stop sampling; keep the samples whose nvidia-smi timestamp falls inside [t_begin, t_end] (time.time())."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = []
        for line in open(self.path):
            p = [x.strip() for x in line.split(',')]
            if len(p) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(p[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                rows.append((ts, float(p[1]), float(p[2]), float(p[3]), p[4:8]))
            except ValueError:
                continue
        os.unlink(self.path)
        inside = [r for r in rows if t_begin is not None and t_begin - 0.05 <= r[0] <= t_end + 0.05]
        window = "timed region"
        if not inside:                          # region shorter than the sampling period: samples under the same load
            inside, window = rows[-3:], "last samples before the end of the timed region"
        sm, mx, pw, reasons = [], [], [], set()
        for _, a, b, c, flags in inside:
            sm.append(a); mx.append(b); pw.append(c)
            for nm, v in zip(names, flags):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        out["window"] = window
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_baseline_serial(inp_full, kind, thetas, budget_s=15.0):
    """The oracle (restatement of the reference) driven the way the reference runs: one walker at a time, 1 core.
    Returns the baseline record, the sample it ran on and the values it produced (re-used by the parity check)."""
    from oracle import lf_oracle
    inp = sample_inputs(inp_full, 1000000)
    n = len(inp['lum'])
    model = lf_oracle.make_model(inp, kind)
    model.lnprob(thetas[0])                                   # warm-up (interp1d set-up, page-in)
    t0 = time.perf_counter()
    vals = []
    for t in thetas[1:]:
        vals.append(model.lnprob(t))
        if time.perf_counter() - t0 > budget_s and len(vals) >= 8:
            break
    dt = time.perf_counter() - t0
    done = len(vals)
    rec = {"value": n * done / dt, "unit": "terms/s", "cores": 1, "kind": "port",
           "sample": "oracle/lf_oracle.py (NumPy restatement, bit-identical to the reference on the golden "
                     "fixtures), serial as the reference runs emcee: %d sources (1/%d of rank 0's shard, every field) "
                     "x %d walkers in %.1f s" % (n, max(1, round(len(inp_full['lum']) / n)), done, dt)}
    return rec, inp, np.asarray(vals, dtype=np.float64)


_POOL_STATE = {}


def _pool_eval(theta):
    return _POOL_STATE['model'].lnprob(theta)


def workload_name(args):
    prec = "FP64" if getattr(args, 'precision', 'f64') == 'f64' else "FP32 loop"
    if args.scaling == 'strong':
        return ("lnprob throughput: %s model, %g sources in total (sharded over the GPUs) x %d walkers, %s "
                "(BASELINE.json configs[1] corner = the north_star target catalogue)" % (KIND_NAMES[args.kind], args.nsources, args.walkers, prec))
    return ("lnprob throughput: %s model, %g sources per GPU x %d walkers, %s (BASELINE.json configs[1] corner; "
            "source-sharded over GPUs as configs[4])" % (KIND_NAMES[args.kind], args.nsources, args.walkers, prec))


def run_reference(args):
    """Reference arm: the reference's own algorithm (NumPy oracle port; the Python reference cannot travel to the
    GPU box) on all host cores, walkers fanned out over a fork pool."""
    import multiprocessing as mp
    from oracle import lf_oracle
    from lumfuncmcmc_b200 import synth
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = min(cores, 64)
    n_sample = 250000
    inp = build_inputs(n_sample, args.kind, seed=1000)
    W = 8 * procs
    thetas = synth.draw_thetas(inp, args.kind, W, seed=7, mode='near', scale=0.02)
    _POOL_STATE['model'] = lf_oracle.make_model(inp, args.kind)
    ctx = mp.get_context('fork')
    with ctx.Pool(procs) as pool:
        for _ in range(args.warmup):
            pool.map(_pool_eval, list(thetas), chunksize=2)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_pool_eval, list(thetas), chunksize=2)
        dt = time.perf_counter() - t0
    value = n_sample * W * args.steps / dt
    n_total = float(args.nsources) * (args.gpus if args.scaling == 'weak' else 1)
    sample = ("%d sources x %d walkers per step (bounded sample of the %g x %d workload), fork Pool(%d) over walkers"
              % (n_sample, W, n_total, args.walkers, procs))
    line = {"impl": "reference", "metric": metric_name(args.kind), "value": value, "unit": "terms/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "sample": sample},
            "cpu_baseline": {"value": value, "unit": "terms/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ensemble_steps": {"value": value / (n_total * args.walkers), "unit": "ensemble steps/s",
                               "note": "derived: the reference's cost is exactly linear in walkers x sources, one ensemble update = "
                                       "one lnprob per walker over %g sources" % n_total},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# engine arm
# ---------------------------------------------------------------------------------------------------------------------
class Bench:
    """Per-process state of the engine arm: ranks, device, collectives."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local_rank))
            dist.barrier()
        from lumfuncmcmc_b200 import synth
        synth.DEVICE = self.local_rank

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device='cuda')
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def global_min(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device='cuda')
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t[0])

    def gather_vectors(self, vec):
        """[vector of rank 0, ..., vector of rank N-1] on every rank (host arrays)."""
        t = self.torch
        d = t.from_numpy(np.ascontiguousarray(vec, dtype=np.float64)).cuda()
        if self.world == 1:
            return [d.cpu().numpy()]
        parts = [t.empty_like(d) for _ in range(self.world)]
        self.dist.all_gather(parts, d)
        return [p.cpu().numpy() for p in parts]

    # -----------------------------------------------------------------------------------------------------------------
    def shard(self, n_total, kind, seed):
        """This rank's shard of an n_total-source catalogue (every rank draws its own part; the quadrature grid is the
        same on every rank: fixed redshift range, luminosity floor = minimum over ALL ranks' sources)."""
        from lumfuncmcmc_b200.dist import shard_bounds
        lo, hi = shard_bounds(int(n_total), self.rank, self.world)
        return build_inputs(hi - lo, kind, seed + self.rank, lum_floor=self.global_min)

    def time_device(self, like, d_th, d_out, steps, warmup):
        t = self.torch
        # warm-up: the W steps asked for (at least 3), then more of the same until about WARMUP_MIN_S of GPU work has run -- a
        # region timed right after a few short steps on a GPU that idled during the set-up has measured 3-7 % slow (2 of 10
        # multi-GPU runs; the regions that followed in the same process were at the expected rate).  The number of extra steps
        # is agreed between the ranks (every step contains a collective).
        n0 = max(3, warmup)
        t0 = time.perf_counter()
        for _ in range(n0):
            like.lnprob_device(d_th, d_out)
        t.cuda.synchronize()
        dt = max(time.perf_counter() - t0, 1e-6)
        extra = int(min(400, max(0, np.ceil((WARMUP_MIN_S - dt) / (dt / n0)))))
        extra = int(self.max_over_ranks([extra])[0])
        for _ in range(extra):
            like.lnprob_device(d_th, d_out)
        self.warmup_steps_done = n0 + extra
        self.sync_all()
        ev0, ev1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        t_begin = time.time()
        ev0.record()
        for _ in range(steps):
            like.lnprob_device(d_th, d_out)
        ev1.record()
        self.sync_all()
        return ev0.elapsed_time(ev1), t_begin, time.time()

    def time_e2e(self, like, thetas, steps):
        for _ in range(3):
            like.lnprob(thetas)
        self.sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = like.lnprob(thetas)
        self.torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, res

    # -----------------------------------------------------------------------------------------------------------------
    def check_parity(self, like, inp, kind, thetas, result, n_oracle_walkers=2, oracle_budget_sources=6.0e7):
        """Correctness of THIS run's numbers, at the size they were measured on.

        exchange      every rank's UN-reduced partial vector is gathered, the vectors are added in rank order on the host,
                      and the sum is compared with what the all-reduce returned (<= 1e-13 relative, -inf sets equal)
        oracle_full   every rank runs the oracle (NumPy restatement of the reference) on its OWN shard for the first
                      walkers; the per-shard log-posteriors add up to the catalogue's (each carries the full quadrature
                      term, so (N - 1) x the zero-source value is taken off); compared with the engine at 1e-10"""
        from oracle import lf_oracle
        t = self.torch
        out = {}
        d_th = t.from_numpy(thetas).cuda()
        partial = like.engine.lnprob_device(d_th)                    # this rank's share, before the exchange
        t.cuda.synchronize()
        parts = self.gather_vectors(partial.cpu().numpy())
        host_sum = parts[0].copy()
        for p in parts[1:]:
            host_sum = host_sum + p
        with np.errstate(invalid='ignore'):
            fin = np.isfinite(host_sum)
            same_inf = bool(np.array_equal(np.isneginf(host_sum), np.isneginf(result)))
            rel = float(np.max(np.abs(result[fin] - host_sum[fin]) / np.abs(host_sum[fin]))) if fin.any() else 0.0
        out["exchange"] = {"max_rel_vs_rank_ordered_host_sum": rel, "inf_sets_equal": same_inf, "tol": 1e-13,
                           "ok": bool(same_inf and rel <= 1e-13)}
        n_loc = len(inp['lum'])
        nw = n_oracle_walkers if n_loc * n_oracle_walkers <= oracle_budget_sources else 1
        th_o = thetas[:nw]
        mine = lf_oracle.lnprob_batch(inp, kind, th_o)
        shards = self.gather_vectors(mine)
        total = np.sum(shards, axis=0)
        if self.world > 1:
            empty = dict(inp)
            for key in ('lum', 'z', 'flux', 'flux_src', 'Om_arr'):
                if key in empty and empty[key] is not None:
                    empty[key] = np.zeros(0)
            empty['field_ind'] = np.zeros(len(inp['field_ind']), dtype=np.int64)
            total = total - (self.world - 1) * lf_oracle.lnprob_batch(empty, kind, th_o)
        with np.errstate(invalid='ignore'):
            fin = np.isfinite(total)
            same_inf = bool(np.array_equal(np.isneginf(total), np.isneginf(result[:nw])))
            rel = float(np.max(np.abs(result[:nw][fin] - total[fin]) / np.abs(total[fin]))) if fin.any() else 0.0
        out["oracle_full_size"] = {"walkers": int(nw), "sources_per_rank": int(n_loc), "max_rel": rel, "inf_sets_equal": same_inf,
                                   "tol": 1e-10, "ok": bool(same_inf and rel <= 1e-10)}
        out["ok"] = bool(out["exchange"]["ok"] and out["oracle_full_size"]["ok"])
        return out

    # -----------------------------------------------------------------------------------------------------------------
    def run_sharded_case(self, name, n_total, W, kind, steps, warmup, seed, mode, scale=0.02, keep=False):
        """One source-sharded throughput case: build this rank's shard, time device-resident and end-to-end, check parity."""
        from lumfuncmcmc_b200 import synth
        from lumfuncmcmc_b200.dist import ShardedLikelihood
        t = self.torch
        args = self.args
        t0 = time.perf_counter()
        inp = self.shard(n_total, kind, seed)
        like = ShardedLikelihood(inp, kind, device=self.local_rank, precision=args.precision, exchange=args.exchange,
                                 wcap=max(4096, W))
        t_setup = time.perf_counter() - t0
        thetas = synth.draw_thetas(inp, kind, W, seed=7, mode=mode, scale=scale)          # same on every rank
        d_th = t.from_numpy(thetas).cuda()
        d_out = t.empty(W, dtype=t.float64, device='cuda')
        ms_dev, t_begin, t_end = self.time_device(like, d_th, d_out, steps, warmup)
        warmup_done = self.warmup_steps_done
        result_dev = d_out.cpu().numpy()
        ms_e2e, result_e2e = self.time_e2e(like, thetas, steps)
        assert np.array_equal(result_e2e, result_dev, equal_nan=True)
        parity = self.check_parity(like, inp, kind, thetas, result_dev)
        # N > 1: a second, identically bracketed region of K steps.  `value` stays the FIRST region (the contract's number);
        # the repeat is reported beside it so that a one-off stall on a shared box (seen once: +7 % on one 4-GPU run, with the
        # host-API loop that followed at the expected rate) is visible as such
        ms_rep = self.time_device(like, d_th, d_out, steps, warmup)[0] if self.world > 1 else ms_dev
        ms_dev, ms_e2e, ms_rep = self.max_over_ranks([ms_dev, ms_e2e, ms_rep])
        terms = float(n_total) * W
        rec = {"workload": name, "sources_total": int(n_total), "sources_per_gpu": len(inp['lum']), "walkers": W, "steps": steps,
               "value": terms * steps / (ms_dev * 1e-3), "unit": "terms/s", "ms_per_step": ms_dev / steps,
               "e2e": {"value": terms * steps / (ms_e2e * 1e-3), "unit": "terms/s", "ms_per_step": ms_e2e / steps,
                       "h2d_bytes_per_step": W * like.ndim * 8, "d2h_bytes_per_step": W * 8},
               "setup_s": t_setup, "parity": parity, "warmup_steps": warmup_done}
        if self.world > 1:
            rec["repeat_region"] = {"ms_per_step": ms_rep / steps, "value": terms * steps / (ms_rep * 1e-3),
                                    "note": "second region of the same K steps, informational; `value` is the first region"}
        if keep:
            return rec, like, inp, thetas, d_th, d_out, result_dev, (t_begin, t_end)
        like.close()
        return rec

    def run_walker_sharded(self, n, W, kind, steps, warmup, mode):
        """Small catalogue: every rank holds all n sources and evaluates W / N walkers; one all-gather of the results."""
        from lumfuncmcmc_b200 import synth
        from lumfuncmcmc_b200.dist import WalkerShardedLikelihood
        from lumfuncmcmc_b200.engine import LikelihoodEngine
        t = self.torch
        inp = build_inputs(n, kind, seed=4300)                          # the same catalogue on every rank
        like = WalkerShardedLikelihood(inp, kind, device=self.local_rank, precision=self.args.precision)
        thetas = synth.draw_thetas(inp, kind, W, seed=11, mode=mode, scale=0.02)
        d_th = t.from_numpy(thetas).cuda()
        d_out = t.empty(W, dtype=t.float64, device='cuda')
        ms_dev, _, _ = self.time_device(like, d_th, d_out, steps, warmup)
        res = d_out.cpu().numpy()
        ms_e2e, res_e2e = self.time_e2e(like, thetas, steps)
        # parity: equal to ONE engine evaluating the whole ensemble (rank 0), and to the oracle for 4 walkers
        ok, rel1, relo = True, 0.0, 0.0
        if self.rank == 0:
            from oracle import lf_oracle
            one = like.engine.lnprob(thetas)
            with np.errstate(invalid='ignore'):
                fin = np.isfinite(one)
                rel1 = float(np.max(np.abs(res[fin] - one[fin]) / np.abs(one[fin])))
                ref = lf_oracle.lnprob_batch(inp, kind, thetas[:4])
                f2 = np.isfinite(ref)
                relo = float(np.max(np.abs(res[:4][f2] - ref[f2]) / np.abs(ref[f2]))) if f2.any() else 0.0
            ok = bool(np.array_equal(np.isneginf(res), np.isneginf(one)) and rel1 <= 1e-13 and relo <= 1e-10
                      and np.array_equal(res_e2e, res, equal_nan=True))
        ms_dev, ms_e2e = self.max_over_ranks([ms_dev, ms_e2e])
        like.close()
        # the device-resident sampler on the same ensemble, walkers sharded over the ranks inside the captured update (the
        # peer-memory exchange gathers the slices)
        like_p = WalkerShardedLikelihood(inp, kind, device=self.local_rank, precision=self.args.precision, exchange='p2p', wcap=W)
        like_p.sampler_run(thetas, 2, seed=3)
        self.sync_all()
        nupd = 10
        run = like_p.sampler_run(thetas, nupd, seed=3)
        ms_upd = self.max_over_ranks([run['device_ms'] / nupd])[0]
        like_p.close()
        terms = float(n) * W
        return {"workload": "walker-sharded: %g sources on every GPU x %d walkers, W / N walkers per GPU, one all-gather of W doubles" % (n, W),
                "device_sampler": {"ms_per_update": ms_upd, "updates_per_s": 1e3 / ms_upd,
                                   "note": "lf_sampler_run with walker sharding: two half-ensemble lnprob calls per update, slices "
                                           "gathered by the peer-memory exchange inside the CUDA graph"},
                "sources": int(n), "walkers": W, "steps": steps, "value": terms * steps / (ms_dev * 1e-3), "unit": "terms/s",
                "ms_per_step": ms_dev / steps,
                "e2e": {"value": terms * steps / (ms_e2e * 1e-3), "unit": "terms/s", "ms_per_step": ms_e2e / steps},
                "parity": {"max_rel_vs_one_engine_full_ensemble": rel1, "max_rel_vs_oracle_4_walkers": relo, "ok": ok}}


def veff_block(eng, inp, n):
    """1/V_eff weights + binned LF + one bootstrap replicate on this GPU's sources (BASELINE.json configs[3]; the HBM-bound
    pass of the path).  The sample is uploaded once (lf_veff_set_sample) and stays resident, as VeffLF uses it."""
    lum_h, flux_h = np.asarray(inp['lum'], dtype=np.float64), np.asarray(eng._flux_host, dtype=np.float64)
    nb = 50
    edges = np.linspace(lum_h.min() * 1.001, lum_h.max(), nb + 1)           # VmaxLumFunc.py:340
    so = float(np.sum(inp['Omega_0']))
    t0 = time.perf_counter()
    eng.veff_set_sample(flux_h, lum_h, inp['field_ind'])
    t_upload = time.perf_counter() - t0
    best_w = best_b = best_call = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        _, cnt_v, _ = eng.veff_bin_resident(inp['Flim'], inp['alpha'], inp['fcmin'], so, 3.0e10, edges)
        best_call = min(best_call, (time.perf_counter() - t0) * 1e3)
        best_w = min(best_w, eng.last_kernel_ms())
    # per-bin counts against NumPy on the host: exact comparisons with the same edges, half-open bins (VmaxLumFunc.py:346-348)
    idx = np.searchsorted(edges, lum_h, side='right') - 1
    want = np.bincount(idx[(idx >= 0) & (idx < nb)], minlength=nb)[:nb]
    per_bin_equal = bool(np.array_equal(cnt_v, want))
    mult = np.bincount(np.random.RandomState(3).randint(n, size=n), minlength=n)  # the reference's resampling (:353)
    for _ in range(3):
        cb, _ = eng.boot_bin(mult)
        best_b = min(best_b, eng.last_kernel_ms())
    inb = (idx >= 0) & (idx < nb)
    want_b = np.bincount(idx[inb], weights=mult[inb], minlength=nb)[:nb]
    boot_equal = bool(np.array_equal(cb, want_b.astype(np.int64)))
    # the same resampling with NumPy's MT19937 stream generated ON the device (VmaxLumFunc.py:353): the first replicate must
    # equal the host-drawn one bin for bin; then the cost of a replicate without the host draw and the 4 N-byte upload
    rs = np.random.RandomState(5)
    eng.boot_mt_set_state(rs.get_state())
    t0 = time.perf_counter()
    cm, _ = eng.boot_bin_mt()
    mult_h = np.bincount(rs.randint(n, size=n), minlength=n)
    want_m = np.bincount(idx[inb], weights=mult_h[inb], minlength=nb)[:nb].astype(np.int64)
    mt_equal = bool(np.array_equal(cm, want_m))
    t0 = time.perf_counter()
    for _ in range(3):
        eng.boot_bin_mt()
    mt_ms = (time.perf_counter() - t0) / 3 * 1e3
    st_dev, st_host = eng.boot_mt_get_state(), None
    for _ in range(3):
        rs.randint(n, size=n)
    st_host = rs.get_state()
    mt_state_equal = bool(st_dev[2] == st_host[2] and np.array_equal(st_dev[1], st_host[1]))
    return {"workload": "1/V_eff weights + binning, %d sources, %d bins, sample resident on the device; one bootstrap replicate" % (n, nb),
            "weights_ms": best_w, "weights_gbs": 26.0 * n / (best_w * 1e-3) / 1e9,
            "replicate_ms": best_b, "replicate_gbs": 14.0 * n / (best_b * 1e-3) / 1e9,
            "algorithmic_bytes_per_source": {"weights": 26, "replicate": 14},
            "e2e_ms_resident_call": best_call, "sample_upload_ms_once": t_upload * 1e3,
            "e2e_note": "host wall clock of lf_veff_bin_resident (kernels + D2H of 50 counts and sums + sync); the per-source "
                        "weights stay on the device and are downloaded only when phifunc is read",
            "mt19937_replicate_ms": mt_ms,
            "mt19937_note": "one bootstrap replicate with np.random.randint(N, size=N)'s own stream drawn on the device (lf_boot_bin_mt), "
                            "host wall clock; counts bit-equal to the host-drawn replicate and generator state equal after 4 replicates: %s / %s"
                            % (mt_equal, mt_state_equal),
            "counts_match_numpy_per_bin": per_bin_equal,
            "bootstrap_counts_match_numpy_per_bin": bool(boot_equal and mt_equal and mt_state_equal)}


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
        return
    import __graft_entry__
    if int(os.environ.get('RANK', '0')) == 0:
        __graft_entry__.build()
    B = Bench(args)
    torch, world, rank, local_rank = B.torch, B.world, B.rank, B.local_rank
    from lumfuncmcmc_b200 import synth
    from lumfuncmcmc_b200.dist import ShardedLikelihood

    W = args.walkers
    n_total = int(args.nsources) * (world if args.scaling == 'weak' else 1)
    mode = 'prior' if args.prior_draws else 'near'
    sampler = ClockSampler(local_rank)         # nvidia-smi takes ~0.1 s to start: launched before the set-up
    sampler.start()
    head, like, inp, thetas, d_th, d_out, result_dev, (t_begin, t_end) = B.run_sharded_case(
        workload_name(args), n_total, W, args.kind, args.steps, args.warmup, seed=1000, mode=mode, keep=True)
    clocks = sampler.stop(t_begin, t_end)
    eng = like.engine
    n = len(inp['lum'])
    info = eng.last_call_info()
    # kernels launched per step: measured over one more device call
    l0 = info['launches']
    like.lnprob_device(d_th, d_out)
    torch.cuda.synchronize()
    launches_per_step = eng.last_call_info()["launches"] - l0
    # live FP64-pipe / MUFU peaks of this GPU (roofline denominators; not in MEASURED_PEAKS.json): best of 3
    peak_dfma = max(eng.fp64_peak(100000)[0] for _ in range(3))
    peak_mufu = max(eng.mufu_peak(100000)[0] for _ in range(3))

    ens = small = compressed = veff = None
    extras = {}
    if not args.no_extras:
        # ---- ensemble steps/s (second half of BASELINE.json's metric): the vectorised stretch-move sampler drives the
        # public API; one step = two half-ensemble lnprob calls + host proposal/accept work.  Every rank runs the same
        # seeded sampler in lockstep (same proposals; the all-reduce inside lnprob keeps them identical).
        from lumfuncmcmc_b200.sampler import EnsembleSampler
        rs = np.random.RandomState(11)
        n_samp_steps = max(3, min(args.steps, 10))
        smp = EnsembleSampler(W, like.ndim, like.lnprob, vectorize=True)
        smp.run_mcmc(thetas, 1, rstate0=rs.get_state())
        B.sync_all()
        t0 = time.perf_counter()
        smp.run_mcmc(smp.chain[:, -1, :], n_samp_steps)
        torch.cuda.synchronize()
        ms_steps = B.max_over_ranks([(time.perf_counter() - t0) * 1e3])[0]
        # the same ensemble on the device-resident sampler (one CUDA graph per update; over several GPUs the per-walker sum
        # runs inside the graph, which needs the peer-memory exchange)
        dev_big = None
        if world == 1 or like.exchange == 'p2p':
            like.sampler_run(thetas, 1, seed=5)
            B.sync_all()
            t0 = time.perf_counter()
            run = like.sampler_run(thetas, n_samp_steps, seed=5)
            t_dev_big = time.perf_counter() - t0
            dev_big = {"steps_per_s": n_samp_steps / t_dev_big, "device_ms_per_step": run['device_ms'] / n_samp_steps,
                       "acceptance": float(np.mean(run['naccepted'])) / n_samp_steps}
        # config-0 size on one GPU (rank 0): 10^4 sources x 100 walkers, launch/host-bound regime
        if rank == 0:
            from lumfuncmcmc_b200.engine import LikelihoodEngine
            from lumfuncmcmc_b200.sampler import DeviceEnsembleSampler
            inp_s = build_inputs(10000, args.kind, seed=4242)
            eng_s = LikelihoodEngine(inp_s, args.kind, device=local_rank)
            th_s = synth.draw_thetas(inp_s, args.kind, 100, seed=9, mode=mode, scale=0.02)
            smp_s = EnsembleSampler(100, eng_s.ndim, eng_s.lnprob, vectorize=True)
            smp_s.run_mcmc(th_s, 20, rstate0=rs.get_state())
            t0 = time.perf_counter()
            smp_s.run_mcmc(smp_s.chain[:, -1, :], 200)
            dt_s = time.perf_counter() - t0
            # the same run with the device-resident sampler (proposals, lnprob, accept and chain on the GPU, one CUDA graph
            # per ensemble update); wall clock includes the upload of the start positions and the download of the chain
            dev_s = DeviceEnsembleSampler(100, eng_s.ndim, eng_s, seed=17)
            dev_s.run_mcmc(th_s, 50)
            dt_d = 1e9
            for _ in range(3):                      # the first long run also pays the first touch of the host chain arrays
                t0 = time.perf_counter()
                dev_s.run_mcmc(dev_s.chain[:, -1, :], 2000)
                dt_d = min(dt_d, time.perf_counter() - t0)
            small = {"workload": "BASELINE.json configs[0] size: 1e4 sources x 100 walkers, 1 GPU",
                     "steps_per_s": 2000 / dt_d, "sampler": "device-resident (lf_sampler_run), best of 3 runs of 2000 updates, wall clock incl. chain D2H",
                     "device_ms_per_step": dev_s.device_ms / 6050, "steps_per_s_device_time": 6050.0e3 / dev_s.device_ms,
                     "acceptance": float(np.mean(dev_s.acceptance_fraction)),
                     "host_sampler": {"steps_per_s": 200 / dt_s, "lnprob_calls_per_s": 400 / dt_s,
                                      "acceptance": float(np.mean(smp_s.acceptance_fraction))}}
            eng_s.close()
        ens = {"value": n_samp_steps / (ms_steps * 1e-3), "unit": "ensemble steps/s",
               "workload": "%d walkers x %d sources in total on %d GPU(s): stretch move, 2 half-ensemble "
                           "lnprob calls per step through the public host API" % (W, n_total, world),
               "steps_timed": n_samp_steps, "device_resident": dev_big, "small": small}

        # ---- opt-in compressed catalogue (not the headline): the same ensemble on weighted pseudo-sources -------------
        if rank == 0 and world == 1 and args.kind == 'free' and args.precision == 'f64':
            t0 = time.perf_counter()
            npseudo = eng.compress_catalogue(eng._flux_host, inp['field_ind'], alpha_max=float(inp.get('alpha_lims', (1.0, 7.0))[1]))
            t_build = time.perf_counter() - t0
            d_out_c = torch.empty(W, dtype=torch.float64, device='cuda')
            for _ in range(3):
                like.lnprob_device(d_th, d_out_c)
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(args.steps):
                like.lnprob_device(d_th, d_out_c)
            c1.record()
            torch.cuda.synchronize()
            ms_c = c0.elapsed_time(c1) / args.steps
            res_c = d_out_c.cpu().numpy()
            fin = np.isfinite(result_dev)
            compressed = {"pseudo_sources": npseudo, "sources": n, "ms_per_step": ms_c,
                          "effective_terms_per_s": float(n) * W / (ms_c * 1e-3), "build_s_host": t_build,
                          "max_rel_diff_vs_brute_force": float(np.max(np.abs(res_c[fin] - result_dev[fin]) / np.abs(result_dev[fin]))),
                          "note": "opt-in (LikelihoodEngine(compress=True)): sum over sources replaced by a weighted sum over "
                                  "Chebyshev pseudo-sources in log10 flux (lumfuncmcmc_b200/compress.py); the quadrature is unchanged "
                                  "and now dominates the call"}
            eng.uncompress_catalogue()

        # ---- 1/V_eff on rank 0's shard.  Never allowed to take the line down: any failure is reported instead. ---------
        if rank == 0 and args.kind == 'free':
            try:
                veff = veff_block(eng, inp, n)
            except Exception as exc:                                                    # pragma: no cover
                veff = {"error": "%s: %s" % (type(exc).__name__, exc)}

    # ---- CPU baseline (rank 0) + its by-product: the oracle's values on a 10^6-source sub-shard, compared with an engine
    # built on exactly that sub-shard -------------------------------------------------------------------------------------
    cpu_rec = sub_parity = None
    if rank == 0 and not args.no_cpu_baseline:
        from lumfuncmcmc_b200.engine import LikelihoodEngine
        cpu_rec, inp_sub, vals = cpu_baseline_serial(inp, args.kind, thetas)
        eng_sub = LikelihoodEngine(inp_sub, args.kind, device=local_rank, precision=args.precision)
        got = eng_sub.lnprob(thetas[1:1 + len(vals)])
        eng_sub.close()
        with np.errstate(invalid='ignore'):
            fin = np.isfinite(vals)
            rel = float(np.max(np.abs(got[fin] - vals[fin]) / np.abs(vals[fin]))) if fin.any() else 0.0
        tol = 1e-10 if args.precision == 'f64' else 1e-5
        sub_parity = {"walkers": int(len(vals)), "sources": int(len(inp_sub['lum'])), "max_rel": rel, "tol": tol,
                      "ok": bool(np.array_equal(np.isneginf(got), np.isneginf(vals)) and rel <= tol)}
    like.close()

    # ---- sub-results at the other configurations (their own catalogues; built, measured and released one at a time) ----
    sub_steps = max(3, min(args.steps, 10))
    if not args.no_extras and args.kind == 'free' and args.precision == 'f64' and not args.prior_draws:
        if world > 1 and args.scaling == 'strong':
            extras["weak"] = B.run_sharded_case("weak scaling: 1e7 sources per GPU x 1024 walkers", 10000000 * world, 1024, 'free',
                                                sub_steps, args.warmup, seed=2000, mode=mode)
        if world > 1:
            extras["config4"] = B.run_sharded_case("BASELINE.json configs[4]: 1e8 sources in total x 2048 walkers, source-sharded",
                                                   100000000, 2048, 'free', max(3, min(args.steps, 5)), args.warmup, seed=3000, mode=mode)
        extras["walker_sharded"] = B.run_walker_sharded(100000, 4096, 'free', sub_steps, args.warmup, mode)
        # BASELINE.json configs[2]: the redshift-evolving model at its own size (k_main<false, Z>), source-sharded like the headline
        extras["config2_z"] = B.run_sharded_case("BASELINE.json configs[2]: redshift-evolving model, 1e6 sources in total x 512 walkers, "
                                                 "source-sharded", 1000000, 512, 'z', sub_steps, args.warmup, seed=5000, mode=mode)

    if rank == 0:
        value, e2e = head["value"], head["e2e"]
        per_gpu = value / world
        step_s = head["ms_per_step"] * 1e-3
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
            hbm_src = 'MEASURED_PEAKS.json'
        except Exception:
            hbm_peak, hbm_src = 6650.0, 'fallback'
        f32 = args.precision == 'f32'
        bytes_per_source = 8 if f32 else BYTES_PER_SOURCE
        alg_bytes = n * bytes_per_source + W * like.ndim * 8 + W * 8
        if f32:
            roof = {"bound": "mufu", "achieved": per_gpu * MUFU_PER_TERM_BY_KIND[args.kind] / 1e12, "peak": peak_mufu / 1e12, "unit": "T MUFU instr/s",
                    "note": "FP32 mode of k_main<false> is SFU bound: achieved = terms/s/GPU x %d MUFU instr/term (rsqrt, lg2, ex2, "
                            "rcp; z model: one ex2); peak = ex2.approx.f32 rate measured live on this GPU (lf_mufu_peak)" % MUFU_PER_TERM_BY_KIND[args.kind]}
        else:
            roof = {"bound": "fp64", "achieved": per_gpu * FP64_INSTR_PER_TERM_BY_KIND[args.kind] * 2 / 1e12, "peak": peak_dfma * 2 / 1e12, "unit": "TFLOP/s",
                    "note": "FP64-FMA-pipe bound kernel k_main<false> (no tensor cores; HBM traffic ~0.02 B/term): achieved = terms/s/GPU "
                            "x %d FP64-pipe instr/term (counted in SASS, tools/sass_loop_mix.py) x 2 FLOP; peak = register-only DFMA "
                            "rate measured live on this GPU (lf_fp64_peak, best of 3: %.3e DFMA/s) x 2 FLOP; the per-walker quadrature (K S^2 points) is "
                            "extra work not counted as terms" % (FP64_INSTR_PER_TERM_BY_KIND[args.kind], peak_dfma)}
        roof["frac"] = roof["achieved"] / roof["peak"]
        if "config2_z" in extras and not f32:
            extras["config2_z"]["frac_of_dfma_peak"] = extras["config2_z"]["value"] / world * FP64_INSTR_PER_TERM_BY_KIND['z'] / peak_dfma
            extras["config2_z"]["frac_note"] = "terms/s/GPU x 9 FP64-pipe instr/term over the DFMA rate measured in this run; the S x S quadrature is not counted"
        # DRAM bytes of one k_main launch from the committed ncu --set full capture of this workload (profiles/), if any
        roof["traffic"] = None
        try:
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'k_main_traffic.json')))
            key = "%s_%s_%d_%d" % (args.kind, args.precision, n, W)
            if key in tr:
                roof["traffic"] = tr[key]["dram_bytes_per_launch"]
                roof["traffic_source"] = tr[key]["source"]
        except Exception:
            pass
        roof["hbm"] = {"achieved": alg_bytes / step_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                       "frac": alg_bytes / step_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                       "algorithmic_bytes_per_step": alg_bytes}
        parity = dict(head["parity"])
        if sub_parity is not None:
            parity["oracle_subshard"] = sub_parity
            parity["ok"] = bool(parity["ok"] and sub_parity["ok"])
        if veff and "counts_match_numpy_per_bin" in veff:
            parity["veff_counts_per_bin"] = {"sources": n, "ok": bool(veff["counts_match_numpy_per_bin"] and veff["bootstrap_counts_match_numpy_per_bin"])}
            parity["ok"] = bool(parity["ok"] and parity["veff_counts_per_bin"]["ok"])
        for key, sub in extras.items():
            parity[key] = sub["parity"]
            parity["ok"] = bool(parity["ok"] and sub["parity"]["ok"])
        line = {
            "metric": metric_name(args.kind, args.precision), "value": value, "unit": "terms/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "warmup_steps_run": head["warmup_steps"], "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": workload_name(args), "kind": args.kind, "sources_total": n_total, "sources_per_gpu": n, "walkers": W,
                       "ndim": like.ndim, "nfields": eng.nfields, "size_ln": eng.size_ln,
                       "walker_draws": mode, "walker_classes_last_step": info,
                       "l2": "source arrays (%.0f MB per GPU) exceed the 126 MB L2; no flush needed" % (n * 16 / 1e6)
                       if n * 16 > 126e6 else "source arrays (%.0f MB per GPU) fit in the 126 MB L2 -- resident by design: every "
                                              "walker group re-sweeps them, and the kernel is FP64-pipe bound, not memory bound" % (n * 16 / 1e6),
                       "parallelism": "sources sharded x%d, walkers replicated, %s of %d B/step" % (
                           world, "NCCL all-reduce" if like.exchange == "nccl" else "peer-memory all-reduce kernel (P2P stores over NVLink)", W * 8)},
            "e2e": dict(e2e, timing="host wall clock around ShardedLikelihood.lnprob (pinned theta H2D, kernels, all-reduce, D2H, sync), max over ranks"),
            "gpu_launches": int(launches_per_step * args.steps),
            "parity": parity,
            "ensemble_steps": ens,
            "clocks": clocks,
            "compressed_catalogue": compressed,
            "veff": veff,
            "roofline": roof,
        }
        line.update(extras)
        if "repeat_region" in head:
            line["repeat_region"] = head["repeat_region"]
        if veff and "weights_gbs" in veff:
            veff["hbm_peak_gbs"] = hbm_peak
            veff["weights_frac_of_hbm"] = veff["weights_gbs"] / hbm_peak
            veff["replicate_frac_of_hbm"] = veff["replicate_gbs"] / hbm_peak
        if cpu_rec is not None:
            line["cpu_baseline"] = cpu_rec
        print(json.dumps(line))
    if world > 1:
        B.dist.barrier()
        B.dist.destroy_process_group()


if __name__ == '__main__':
    main()
