#!/usr/bin/env python
"""bench.py -- walker x source lnL terms/s of the batched lnprob on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference algorithm on the host cores)

Workload (BASELINE.json configs[1], the corner the target is quoted on): free-completeness single-z model
(ndim 9, K = 5 fields, S = 101), 10^7 synthetic sources PER GPU x 1024 walkers, FP64.  A "step" is one batched
lnprob call over the whole ensemble (what a vectorised emcee hands over per half-step, here the full ensemble).
At N GPUs sources are sharded (weak scaling: 10^7 per GPU), walkers replicated, the quadrature split by walker,
one NCCL all-reduce of W doubles per step.

value  : terms/s, inputs resident in HBM (theta on the device), CUDA events on the launching stream, max over ranks.
e2e    : same metric through the public host API (ShardedLikelihood.lnprob: pinned-host theta -> H2D -> kernels
         -> all-reduce -> D2H of W doubles -> sync), host wall clock, max over ranks.
roofline: the loop is FP64-FMA-pipe bound (no tensor cores, HBM traffic ~0.02 B/term): achieved = terms/s/GPU x 23
         FP64-pipe instructions per term (counted in the SASS of k_main<false>) x 2 FLOP, against the register-only
         DFMA rate measured live on the same GPU (lf_fp64_peak) x 2 FLOP.  HBM figures are reported beside it.
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

FP64_INSTR_PER_TERM = 23          # DFMA/DADD/DMUL per (walker, source) term in k_main<false, FREE> (tools/sass_loop_mix.py)
FP64_INSTR_PER_TERM_BY_KIND = {'free': 23, 'z': 9, 'fixed': 0}    # fixed: the source sum is sufficient statistics (quadrature only)
MUFU_PER_TERM_BY_KIND = {'free': 4, 'z': 1, 'fixed': 0}
MUFU_PER_TERM = 4                 # rsqrt, lg2, ex2, rcp per term in the FP32 mode of the loop
BYTES_PER_SOURCE = 16             # (log10 flux, flux) per source per sweep
METRIC = "walker x source lnL terms/sec (batched lnprob, free-completeness single-z, FP64)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='engine', choices=['engine', 'reference'])
    ap.add_argument('--nsources', type=float, default=1.0e7, help='sources per GPU')
    ap.add_argument('--walkers', type=int, default=1024)
    ap.add_argument('--kind', default='free', choices=['free', 'fixed', 'z'])
    ap.add_argument('--precision', default='f64', choices=['f64', 'f32'], help='arithmetic of the walker x source loop')
    ap.add_argument('--exchange', default='nccl', choices=['nccl', 'p2p'],
                    help="multi-GPU sum of the per-walker partials: NCCL all-reduce or the engine's peer-memory kernel")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--prior-draws', action='store_true', help='walkers ~ U(prior) instead of a converged ensemble')
    return ap.parse_args()


def build_inputs(n, kind, seed):
    from lumfuncmcmc_b200 import synth
    cat = synth.make_catalogue(n, seed=seed, evolve=(0.3, -0.2) if kind == 'z' else None)
    return synth.direct_inputs(cat, nknots=4096, size_ln=101 if kind == 'free' else 201, tabulated=(kind != 'free'))


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
    """The oracle (restatement of the reference) driven the way the reference runs: one walker at a time, 1 core."""
    from oracle import lf_oracle
    inp = sample_inputs(inp_full, 1000000)
    n = len(inp['lum'])
    model = lf_oracle.make_model(inp, kind)
    model.lnprob(thetas[0])                                   # warm-up (interp1d set-up, page-in)
    t0 = time.perf_counter()
    done = 0
    for t in thetas[1:]:
        model.lnprob(t)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": n * done / dt, "unit": "terms/s", "cores": 1, "kind": "port",
            "sample": "oracle/lf_oracle.py (NumPy restatement, bit-identical to the reference on the golden "
                      "fixtures), serial as the reference runs emcee: %d sources (1/%d of rank 0's shard, every field) "
                      "x %d walkers in %.1f s" % (n, max(1, round(len(inp_full['lum']) / n)), done, dt)}


_POOL_STATE = {}


def _pool_eval(theta):
    return _POOL_STATE['model'].lnprob(theta)


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
    sample = ("%d sources x %d walkers per step (bounded sample of the %g x %d workload), fork Pool(%d) over walkers"
              % (n_sample, W, args.nsources, args.walkers, procs))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "terms/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "sample": sample},
            "cpu_baseline": {"value": value, "unit": "terms/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ensemble_steps": {"value": value / (float(args.nsources) * args.gpus * args.walkers), "unit": "ensemble steps/s",
                               "note": "derived: the reference's cost is exactly linear in walkers x sources, one ensemble update = "
                                       "one lnprob per walker over %g x %d sources" % (args.nsources, args.gpus)},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_name(args):
    return ("lnprob throughput: %s-completeness model, %g sources per GPU x %d walkers, %s "
            "(BASELINE.json configs[1]; source-sharded over GPUs as configs[4])" %
            (args.kind, args.nsources, args.walkers, "FP64" if getattr(args, 'precision', 'f64') == 'f64' else "FP32 loop"))


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    import __graft_entry__
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if rank == 0:
        __graft_entry__.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        dist.barrier()
    from lumfuncmcmc_b200 import synth
    from lumfuncmcmc_b200.dist import ShardedLikelihood

    n = int(args.nsources)
    W = args.walkers
    inp = build_inputs(n, args.kind, seed=1000 + rank)        # this rank's shard
    like = ShardedLikelihood(inp, args.kind, device=local_rank, precision=args.precision, exchange=args.exchange,
                             wcap=max(4096, W))
    eng = like.engine
    mode = 'prior' if args.prior_draws else 'near'
    thetas = synth.draw_thetas(inp, args.kind, W, seed=7, mode=mode, scale=0.02)     # same on every rank
    d_th = torch.from_numpy(thetas).cuda()
    d_out = torch.empty(W, dtype=torch.float64, device='cuda')

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # live FP64-pipe peak of this GPU (roofline denominator; not in MEASURED_PEAKS.json)
    peak_dfma, _ = eng.fp64_peak(100000)
    peak_mufu, _ = eng.mufu_peak(100000)

    # ---- device-resident timing ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)         # nvidia-smi takes ~0.1 s to start: launched before the warm-up
    sampler.start()
    for _ in range(max(3, args.warmup)):
        like.lnprob_device(d_th, d_out)
    sync_all()
    launches0 = eng.last_call_info()['launches']
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_begin = time.time()
    ev0.record()
    for _ in range(args.steps):
        like.lnprob_device(d_th, d_out)
    ev1.record()
    sync_all()
    clocks = sampler.stop(t_begin, time.time())
    ms_dev = ev0.elapsed_time(ev1)
    info = eng.last_call_info()
    launches = info['launches'] - launches0
    result_dev = d_out.cpu().numpy()

    # ---- end to end through the public host API ----------------------------------------------------
    for _ in range(2):
        like.lnprob(thetas)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        result_e2e = like.lnprob(thetas)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert np.array_equal(result_e2e, result_dev, equal_nan=True)

    # ---- ensemble steps/s (second half of BASELINE.json's metric): the vectorised stretch-move sampler drives the
    # public API; one step = two half-ensemble lnprob calls + host proposal/accept work.  Every rank runs the same
    # seeded sampler in lockstep (same proposals; the all-reduce inside lnprob keeps them identical).
    from lumfuncmcmc_b200.sampler import EnsembleSampler
    rs = np.random.RandomState(11)
    n_samp_steps = max(3, min(args.steps, 10))
    smp = EnsembleSampler(W, like.ndim, like.lnprob, vectorize=True)
    smp.run_mcmc(thetas, 1, rstate0=rs.get_state())
    sync_all()
    t0 = time.perf_counter()
    smp.run_mcmc(smp.chain[:, -1, :], n_samp_steps)
    torch.cuda.synchronize()
    t_steps = time.perf_counter() - t0
    # the same ensemble on the device-resident sampler (one CUDA graph per update; over several GPUs the per-walker sum runs
    # inside the graph, which needs the peer-memory exchange)
    dev_big = None
    if world == 1 or like.exchange == 'p2p':
        like.sampler_run(thetas, 1, seed=5)
        sync_all()
        t0 = time.perf_counter()
        run = like.sampler_run(thetas, n_samp_steps, seed=5)
        t_dev_big = time.perf_counter() - t0
        dev_big = {"steps_per_s": n_samp_steps / t_dev_big, "device_ms_per_step": run['device_ms'] / n_samp_steps,
                   "acceptance": float(np.mean(run['naccepted'])) / n_samp_steps}
    # config-1 size on one GPU (rank 0): 10^4 sources x 100 walkers, launch/host-bound regime
    small = None
    if rank == 0:
        inp_s = build_inputs(10000, args.kind, seed=4242)
        like_s = ShardedLikelihood.__new__(ShardedLikelihood)
        from lumfuncmcmc_b200.engine import LikelihoodEngine
        eng_s = LikelihoodEngine(inp_s, args.kind, device=local_rank)
        th_s = synth.draw_thetas(inp_s, args.kind, 100, seed=9, mode=mode, scale=0.02)
        smp_s = EnsembleSampler(100, eng_s.ndim, eng_s.lnprob, vectorize=True)
        smp_s.run_mcmc(th_s, 20, rstate0=rs.get_state())
        t0 = time.perf_counter()
        smp_s.run_mcmc(smp_s.chain[:, -1, :], 200)
        dt_s = time.perf_counter() - t0
        # the same run with the device-resident sampler (proposals, lnprob, accept and chain on the GPU, one CUDA graph
        # per ensemble update); wall clock includes the upload of the start positions and the download of the chain
        from lumfuncmcmc_b200.sampler import DeviceEnsembleSampler
        dev_s = DeviceEnsembleSampler(100, eng_s.ndim, eng_s, seed=17)
        dev_s.run_mcmc(th_s, 50)
        dt_d = 1e9
        for _ in range(3):                      # the first long run also pays the first touch of the host chain arrays
            t0 = time.perf_counter()
            dev_s.run_mcmc(dev_s.chain[:, -1, :], 2000)
            dt_d = min(dt_d, time.perf_counter() - t0)
        small = {"workload": "BASELINE.json configs[0] size: 1e4 sources x 100 walkers, 1 GPU",
                 "steps_per_s": 2000 / dt_d, "sampler": "device-resident (lf_sampler_run), best of 3 runs of 2000 updates, wall clock incl. chain D2H",
                 "device_ms_per_step": dev_s.device_ms / 6050, "steps_per_s_device_time": 6050.0e3 / dev_s.device_ms, "acceptance": float(np.mean(dev_s.acceptance_fraction)),
                 "host_sampler": {"steps_per_s": 200 / dt_s, "lnprob_calls_per_s": 400 / dt_s,
                                  "acceptance": float(np.mean(smp_s.acceptance_fraction))}}
        eng_s.close()

    # ---- opt-in compressed catalogue (not the headline): the same ensemble on weighted pseudo-sources ------------------
    compressed = None
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

    # ---- 1/V_eff weights + binned LF + one bootstrap replicate on this GPU's sources (BASELINE.json configs[3]; the
    # HBM-bound pass of the path; not the headline).  Never allowed to take the line down: any failure is reported instead.
    veff = None
    if rank == 0 and args.kind == 'free':
        try:
            lum_h, flux_h = np.asarray(inp['lum'], dtype=np.float64), np.asarray(eng._flux_host, dtype=np.float64)
            nb = 50
            edges = np.linspace(lum_h.min() * 1.001, lum_h.max(), nb + 1)           # VmaxLumFunc.py:340
            best_w = best_b = 1e9
            for _ in range(3):
                _, cnt_v, _ = eng.veff_bin(flux_h, lum_h, inp['field_ind'], inp['Flim'], inp['alpha'], inp['fcmin'],
                                           float(np.sum(inp['Omega_0'])), 3.0e10, edges, want_phi=False)
                best_w = min(best_w, eng.last_kernel_ms())
            in_range = int(np.count_nonzero((lum_h >= edges[0]) & (lum_h < edges[-1])))
            mult = np.bincount(np.random.RandomState(3).randint(n, size=n), minlength=n)  # the reference's resampling (:353)
            for _ in range(3):
                eng.boot_bin(mult)
                best_b = min(best_b, eng.last_kernel_ms())
            veff = {"workload": "1/V_eff weights + binning, %d sources, %d bins; one bootstrap replicate on the resident sample" % (n, nb),
                    "weights_ms": best_w, "weights_gbs": 26.0 * n / (best_w * 1e-3) / 1e9,
                    "replicate_ms": best_b, "replicate_gbs": 14.0 * n / (best_b * 1e-3) / 1e9,
                    "algorithmic_bytes_per_source": {"weights": 26, "replicate": 14},
                    "counts_match_numpy": bool(int(cnt_v.sum()) == in_range)}
        except Exception as exc:                                                    # pragma: no cover
            veff = {"error": "%s: %s" % (type(exc).__name__, exc)}

    t = torch.tensor([ms_dev, t_e2e * 1e3, t_steps * 1e3], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_steps = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        terms_per_step = float(n) * world * W
        value = terms_per_step * args.steps / (ms_dev * 1e-3)
        e2e_value = terms_per_step * args.steps / (ms_e2e * 1e-3)
        per_gpu = value / world
        step_s = ms_dev * 1e-3 / args.steps
        hbm_peak = 6452.8
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
                            "rate measured live on this GPU (lf_fp64_peak: %.3e DFMA/s) x 2 FLOP; the per-walker quadrature (K S^2 points) is "
                            "extra work not counted as terms" % (FP64_INSTR_PER_TERM_BY_KIND[args.kind], peak_dfma)}
        roof["frac"] = roof["achieved"] / roof["peak"]
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
        line = {
            "metric": METRIC if not f32 else METRIC.replace("FP64", "FP32 loop"), "value": value, "unit": "terms/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": workload_name(args), "kind": args.kind, "sources_per_gpu": n, "walkers": W,
                       "ndim": like.ndim, "nfields": eng.nfields, "size_ln": eng.size_ln,
                       "walker_draws": mode, "walker_classes_last_step": info,
                       "l2": "source arrays (%.0f MB per GPU) exceed the 126 MB L2; no flush needed" % (n * 16 / 1e6)
                       if n * 16 > 126e6 else "inputs fit in L2 (resident by design: re-swept by every walker group)",
                       "parallelism": "sources sharded x%d, walkers replicated, %s of %d B/step" % (
                           world, "NCCL all-reduce" if like.exchange == "nccl" else "peer-memory all-reduce kernel (P2P stores over NVLink)", W * 8)},
            "e2e": {"value": e2e_value, "unit": "terms/s", "h2d_bytes_per_step": W * like.ndim * 8,
                    "d2h_bytes_per_step": W * 8, "ms_per_step": ms_e2e / args.steps,
                    "timing": "host wall clock around ShardedLikelihood.lnprob (pinned theta H2D, kernels, all-reduce, D2H, sync), max over ranks"},
            "gpu_launches": int(launches),
            "ensemble_steps": {"value": n_samp_steps / (ms_steps * 1e-3), "unit": "ensemble steps/s",
                               "workload": "%d walkers x %g sources per GPU x %d GPU(s): stretch move, 2 half-ensemble "
                                           "lnprob calls per step through the public host API" % (W, args.nsources, world),
                               "steps_timed": n_samp_steps, "device_resident": dev_big, "small": small},
            "clocks": clocks,
            "compressed_catalogue": compressed,
            "veff": veff,
            "roofline": roof,
        }
        if veff and "weights_gbs" in veff:
            veff["hbm_peak_gbs"] = hbm_peak
            veff["weights_frac_of_hbm"] = veff["weights_gbs"] / hbm_peak
            veff["replicate_frac_of_hbm"] = veff["replicate_gbs"] / hbm_peak
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_serial(inp, args.kind, thetas)
        print(json.dumps(line))
    like.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
