"""Source-sharded likelihood across GPUs: one process per GPU, one exchange of W doubles per call.

``lnprob`` is a plain sum over independent sources (reference lumfuncmcmc.py:370, :388; lumfuncmcmc_z.py:371),
so rank r holds its own shard of sources, evaluates the shard's per-walker partial log-likelihood and the
quadrature term of the walkers ``w % world == r``, and a single ``all_reduce(SUM)`` over the W-vector gives
every rank the full log-posterior (-inf propagates through the sum).  The message is W*8 bytes (8 KiB at
W = 1024): latency-bound on NVLink/NVSwitch, so NCCL's small-message path is the right tool and there is no
bandwidth-bound exchange to fuse into the compute kernel.

``exchange='p2p'`` replaces the NCCL call by the engine's own peer-memory kernel (``lf_allreduce_device``): every rank
stores its W partials into a slot of every peer's buffer over NVLink, raises a flag, waits for the peers' flags and
adds the slots in rank order -- identical bits on every rank, no second library between ``k_finish`` and the caller,
capturable in a CUDA graph.  ``torch.distributed`` is then used only once, to hand round the CUDA IPC handles.

With ``torch.distributed`` not initialised (or world size 1) this degrades to the single-GPU engine.
The host-side logic (sharding, share assignment, reduction semantics) is exercised on CPU with the gloo
backend in tests/test_dist_cpu.py through :func:`reduce_partials`.
"""
import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous split of n items: rank r gets [lo, hi)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_inputs(inp, rank, world):
    """Rank r's share of every field (field membership travels as the shard's own field_ind)."""
    fi = np.asarray(inp['field_ind'], dtype=np.int64)
    keep, new_fi = [], [0]
    for k in range(len(fi) - 1):
        lo, hi = shard_bounds(int(fi[k + 1] - fi[k]), rank, world)
        keep.append(np.arange(fi[k] + lo, fi[k] + hi))
        new_fi.append(new_fi[-1] + (hi - lo))
    keep = np.concatenate(keep) if keep else np.zeros(0, dtype=np.int64)
    out = dict(inp)
    for key in ('lum', 'z', 'Om_arr', 'flux', 'flux_src'):
        if key in inp and inp[key] is not None:
            out[key] = np.asarray(inp[key])[keep]
    out['field_ind'] = np.array(new_fi, dtype=np.int64)
    return out


def reduce_partials(partial, group=None):
    """In-place SUM all-reduce of a per-walker partial lnprob tensor (torch, any backend)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def check_peer_geometry(gathered, world, wcap):
    """``gathered[r] = (handle, rank, world, wcap)`` as every rank reported it; raises unless all agree with this rank."""
    for r, g in enumerate(gathered):
        if g[1] != r or g[2] != world or g[3] != wcap:
            raise RuntimeError("peer buffers disagree: rank %d reports (rank=%d, world=%d, wcap=%d), expected (%d, %d, %d)"
                               % (r, g[1], g[2], g[3], r, world, wcap))


def walker_slice(W, rank, world):
    """Walkers [lo, hi) of a W-walker ensemble that rank r evaluates under walker sharding (contiguous, balanced)."""
    return shard_bounds(W, rank, world)


def gather_walker_results(local, W, group=None):
    """All-gather of the per-rank result slices back into the full (W,) vector, every rank gets all of it."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [walker_slice(W, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.full((width,), float('nan'), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)])


class WalkerShardedLikelihood:
    """Small catalogues: every rank holds ALL sources and evaluates its slice of the walkers; the only
    exchange is the all-gather of W/world results -- no sum over ranks.  Every walker's value is computed by one GPU
    exactly as a single-GPU call on that slice would (the slab partition of the sources depends on the number of walkers
    in the call, so the bits can differ from a full-ensemble call in the last place, ~1e-16 relative)."""

    def __init__(self, inp, kind, device=None, group=None, precision='f64', exchange='nccl', wcap=4096):
        import torch
        import torch.distributed as dist
        from .engine import LikelihoodEngine
        self.torch, self.group = torch, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.engine = LikelihoodEngine(inp, kind, device=self.device, precision=precision)
        self.ndim = self.engine.ndim
        if exchange not in ('nccl', 'p2p'):
            raise ValueError("exchange must be 'nccl' or 'p2p'")
        # 'p2p': the engine's peer-memory kernel gathers the slices (each rank contributes zeros outside its own), which is
        # what lets the device-resident sampler run walker-sharded inside its captured graphs (sampler_run)
        self.exchange = exchange if self.world > 1 else 'nccl'
        self.wcap = int(wcap)
        if self.exchange == 'p2p':
            handle = self.engine.peer_buffer_create(self.rank, self.world, self.wcap)
            gathered = [None] * self.world
            dist.all_gather_object(gathered, (handle, self.rank, self.world, self.wcap), group=group)
            check_peer_geometry(gathered, self.world, self.wcap)
            self.engine.peer_buffer_connect([g[0] for g in gathered])
            self.engine.set_walker_sharding(True)
            dist.barrier(group=group)

    def sampler_run(self, pos0, nsteps, seed, a=2.0, step0=0):
        """Device-resident ensemble run with the walkers of every half-ensemble sharded over the ranks (all ranks get the
        same chain).  More than one rank needs ``exchange='p2p'``."""
        if self.world > 1 and self.exchange != 'p2p':
            raise RuntimeError("the walker-sharded device-resident sampler needs exchange='p2p'")
        out = self.engine.sampler_run(pos0, nsteps, seed, a=a, step0=step0)
        if self.exchange == 'p2p' and self.engine.peer_timed_out():
            raise RuntimeError("peer-memory all-reduce timed out waiting for another rank")
        return out

    def lnprob_device(self, d_thetas, d_out=None):
        """Device-resident call on torch's current stream: this rank's slice of ``d_thetas`` (all W rows, identical on
        every rank) -> kernels -> all-gather; returns the full (W,) vector on the device."""
        t = self.torch
        W = d_thetas.shape[0]
        lo, hi = walker_slice(W, self.rank, self.world)
        if self.world == 1:
            return self.engine.lnprob_device(d_thetas, d_out)
        if self.exchange == 'p2p':
            if self.engine.peer_timed_out():
                raise RuntimeError("peer-memory all-reduce timed out waiting for another rank; results since then are NaN")
            if W > self.wcap:
                raise ValueError("more walkers (%d) than the peer buffers hold (wcap=%d)" % (W, self.wcap))
            full = t.zeros(W, dtype=t.float64, device=d_thetas.device) if d_out is None else d_out.zero_()
            if hi > lo:
                self.engine.lnprob_device(d_thetas[lo:hi], full[lo:hi])
            return self.engine.allreduce_device(full)          # x + 0 = x: the rank-ordered sum is the all-gather
        local = self.engine.lnprob_device(d_thetas[lo:hi]) if hi > lo else t.zeros(0, dtype=t.float64, device=d_thetas.device)
        full = gather_walker_results(local, W, self.group)
        if d_out is not None:
            d_out.copy_(full)
            return d_out
        return full

    def lnprob(self, thetas):
        t = self.torch
        th = np.ascontiguousarray(np.atleast_2d(np.asarray(thetas, dtype=np.float64)))
        W = th.shape[0]
        lo, hi = walker_slice(W, self.rank, self.world)
        if self.world > 1 and self.exchange == 'p2p':
            with t.cuda.device(self.device):
                d_th = t.from_numpy(th).to(t.device('cuda', self.device))
                out = self.lnprob_device(d_th).cpu().numpy()
            if self.engine.peer_timed_out():
                raise RuntimeError("peer-memory all-reduce timed out waiting for another rank")
            return out
        mine = self.engine.lnprob(th[lo:hi]) if hi > lo else np.zeros(0)
        if self.world == 1:
            return mine
        local = t.from_numpy(np.ascontiguousarray(mine)).to(t.device('cuda', self.device))
        return gather_walker_results(local, W, self.group).cpu().numpy()

    def close(self):
        self.engine.close()


class ShardedLikelihood:
    """Public multi-GPU API: ``lnprob(thetas_host) -> lnprob_host`` with H2D, kernels, all-reduce, D2H.

    ``inp`` is THIS rank's shard (already split, e.g. by :func:`shard_inputs`)."""

    def __init__(self, inp, kind, device=None, group=None, precision='f64', exchange='nccl', wcap=4096, compress=False):
        import torch
        import torch.distributed as dist
        from .engine import LikelihoodEngine
        self.torch = torch
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.engine = LikelihoodEngine(inp, kind, device=self.device, quadrature_share=(self.rank, self.world),
                                       precision=precision, compress=compress)
        self.ndim = self.engine.ndim
        self._cap = 0
        if exchange not in ('nccl', 'p2p'):
            raise ValueError("exchange must be 'nccl' or 'p2p'")
        self.exchange = exchange if self.world > 1 else 'nccl'
        self.wcap = int(wcap)
        if self.exchange == 'p2p':
            handle = self.engine.peer_buffer_create(self.rank, self.world, self.wcap)
            # the slot offsets inside every rank's buffer are computed from (world, wcap): a rank that created its buffer
            # with other values would receive stores in the wrong place, so the geometry travels with the handle
            mine = (handle, self.rank, self.world, self.wcap)
            gathered = [None] * self.world
            dist.all_gather_object(gathered, mine, group=group)
            check_peer_geometry(gathered, self.world, self.wcap)
            self.engine.peer_buffer_connect([g[0] for g in gathered])
            dist.barrier(group=group)

    def _ensure(self, W):
        if W <= self._cap:
            return
        t = self.torch
        dev = t.device('cuda', self.device)
        self._h_th = t.empty((W, self.ndim), dtype=t.float64).pin_memory()
        self._h_out = t.empty(W, dtype=t.float64).pin_memory()
        self._d_th = t.empty((W, self.ndim), dtype=t.float64, device=dev)
        self._d_out = t.empty(W, dtype=t.float64, device=dev)
        self._cap = W

    def lnprob_device(self, d_thetas, d_out=None):
        """Device-resident call on torch's current stream: kernels + all-reduce, asynchronous.  A peer time-out of an
        earlier call (sticky; its result and all later ones are NaN) raises here."""
        if self.exchange == 'p2p' and self.engine.peer_timed_out():
            raise RuntimeError("peer-memory all-reduce timed out waiting for another rank; results since then are NaN")
        d_out = self.engine.lnprob_device(d_thetas, d_out)
        if self.exchange == 'p2p':
            if d_out.shape[0] > self.wcap:
                raise ValueError("more walkers (%d) than the peer buffers hold (wcap=%d)" % (d_out.shape[0], self.wcap))
            return self.engine.allreduce_device(d_out)
        return reduce_partials(d_out, self.group)

    def lnprob(self, thetas):
        t = self.torch
        th = np.ascontiguousarray(np.atleast_2d(np.asarray(thetas, dtype=np.float64)))
        W = th.shape[0]
        if self.world == 1:
            # one rank: the engine's own host entry point (pinned staging, H2D, kernels, D2H, sync inside lf_lnprob_batch)
            return self.engine.lnprob(th)
        self._ensure(W)
        self._h_th[:W].copy_(t.from_numpy(th))
        with t.cuda.device(self.device):
            self._d_th[:W].copy_(self._h_th[:W], non_blocking=True)
            out = self.lnprob_device(self._d_th[:W], self._d_out[:W])
            self._h_out[:W].copy_(out, non_blocking=True)
            t.cuda.current_stream().synchronize()
        if self.exchange == 'p2p' and self.engine.peer_timed_out():
            raise RuntimeError("peer-memory all-reduce timed out waiting for another rank")
        return self._h_out[:W].numpy().copy()

    def sampler_run(self, pos0, nsteps, seed, a=2.0, step0=0):
        """Device-resident ensemble run over all ranks (every rank gets the same chain).  Needs ``exchange='p2p'`` when
        there is more than one rank: the sum over ranks runs inside the captured CUDA graph of each update."""
        if self.world > 1 and self.exchange != 'p2p':
            raise RuntimeError("the device-resident sampler over several GPUs needs exchange='p2p'")
        out = self.engine.sampler_run(pos0, nsteps, seed, a=a, step0=step0)
        if self.exchange == 'p2p' and self.engine.peer_timed_out():
            raise RuntimeError("peer-memory all-reduce timed out waiting for another rank")
        return out

    def close(self):
        self.engine.close()
