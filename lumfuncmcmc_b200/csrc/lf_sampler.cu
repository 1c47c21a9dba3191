// lf_sampler.cu -- device-resident affine-invariant ensemble sampler (replaces emcee.EnsembleSampler.run_mcmc,
// reference lumfuncmcmc.py:489-491) behind include/lf_engine.h.
#include "lf_internal.cuh"

// ------------------------------------------------------------------------------------------------
// device-resident ensemble sampler (stretch move)
// ------------------------------------------------------------------------------------------------
struct SamplerArgs {
    int W, half, ndim;
    double a;
    uint32_t k0, k1;
    const long long* step;      // device counter: index of the current ensemble update
    long long step0;            // value of *step at the first replay (chain rows are relative to it)
    double* pos; double* lp;    // [W][ndim], [W]
    double* prop; double* lpnew; double* lnz; double* lnu;     // [half][ndim], [half] x 3
    double* lp_next;            // [W] second log-posterior buffer: steps with odd index read lp_next and write lp
    double* chain; double* lnp; long long* nacc;               // [nsteps][W][ndim], [nsteps][W], [W]  (chain / lnp may be NULL)
};

// proposals for the walkers of half h (h = 0: [0, half), h = 1: [half, W)) against the other half.
// One thread per (walker, parameter): the Philox call and z are recomputed per parameter (cheap), the copies run in parallel.
__global__ void k_stretch_propose(SamplerArgs s, int h) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= s.half * s.ndim) return;
    const int i = idx / s.ndim, d = idx - i * s.ndim;
    const int me = h * s.half + i;
    const long long step = *s.step;
    uint32_t r[4];
    philox4x32_10((uint32_t)me, (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)h, s.k0, s.k1, r);
    // z ~ g(z) propto 1/sqrt(z) on [1/a, a]:  z = ((a - 1) u + 1)^2 / a      (rounded operation by operation: the host
    // replay in tests/ reproduces the chain bit for bit)
    const double t = __dadd_rn(__dmul_rn(s.a - 1.0, u01(r[0])), 1.0);
    const double z = __ddiv_rn(__dmul_rn(t, t), s.a);
    const int partner = (1 - h) * s.half + (int)(((unsigned long long)r[1] * (unsigned long long)s.half) >> 32);
    const double pp = s.pos[(long long)partner * s.ndim + d], pm = s.pos[(long long)me * s.ndim + d];
    s.prop[(long long)i * s.ndim + d] = __dsub_rn(pp, __dmul_rn(__dsub_rn(pp, pm), z));
    if (d == 0) {
        s.lnz[i] = (s.ndim - 1.0) * log(z);
        s.lnu[i] = log(u01(r[2]));
    }
}

// accept / reject and chain write, one thread per (walker, parameter).  The log-posteriors ping-pong between two buffers
// by step parity (every thread of a walker reads the OLD value from one buffer, the d == 0 thread writes the other), so
// no thread can see a half-updated state.
__global__ void k_stretch_accept(SamplerArgs s, int h) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= s.half * s.ndim) return;
    const int i = idx / s.ndim, d = idx - i * s.ndim;
    const int me = h * s.half + i;
    const long long step = *s.step;
    const double* lp_cur = (step & 1) ? s.lp_next : s.lp;
    double* lp_out = (step & 1) ? s.lp : s.lp_next;
    const double lp_old = lp_cur[me], lp_new = s.lpnew[i];
    const double lnratio = s.lnz[i] + lp_new - lp_old;              // NaN (inf - inf) compares false: rejected
    const bool acc = s.lnu[i] < lnratio;
    const long long row = step - s.step0;
    const double v = acc ? s.prop[(long long)i * s.ndim + d] : s.pos[(long long)me * s.ndim + d];
    if (acc) s.pos[(long long)me * s.ndim + d] = v;
    if (s.chain) s.chain[(row * s.W + me) * s.ndim + d] = v;
    if (d == 0) {
        const double lp_keep = acc ? lp_new : lp_old;
        lp_out[me] = lp_keep;
        if (acc) s.nacc[me] += 1;
        if (s.lnp) s.lnp[row * s.W + me] = lp_keep;
    }
}
__global__ void k_step_advance(long long* step) { *step += 1; }

extern "C" int lf_sampler_run(lf_ctx* c, const double* pos0, int64_t W, int64_t nsteps, uint64_t seed, double a, int64_t step0,
                              double* chain, double* lnprob, int64_t* naccepted, double* pos_out, double* lnprob_out) {
    if (!c || !pos0) return fail("lf_sampler_run: null argument");
    if (W < 2 || (W & 1)) return fail("lf_sampler_run: the number of walkers must be even");
    if (nsteps < 0 || !(a > 1.0)) return fail("lf_sampler_run: need nsteps >= 0 and a > 1");
    if (!c->have_sources || !c->have_grid) return fail("lf_sampler_run: call lf_set_sources and lf_set_grid first");
    CK(cudaSetDevice(c->device));
    const int ndim = c->ndim, half = (int)(W / 2);
    cudaStream_t st = c->stream;
    double *d_pos = nullptr, *d_lp = nullptr, *d_prop = nullptr, *d_lpnew = nullptr, *d_lnz = nullptr, *d_lnu = nullptr;
    double *d_chain = nullptr, *d_lnp = nullptr, *d_lpnext = nullptr;
    long long *d_nacc = nullptr, *d_step = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int rc = 1;
    auto cleanup = [&]() {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        dfree(d_pos); dfree(d_lp); dfree(d_prop); dfree(d_lpnew); dfree(d_lnz); dfree(d_lnu);
        dfree(d_chain); dfree(d_lnp); dfree(d_nacc); dfree(d_step); dfree(d_lpnext);
        return rc;
    };
#define SCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return cleanup();                                                                       \
        }                                                                                           \
    } while (0)
    SCK(cudaMalloc(&d_pos, sizeof(double) * W * ndim));
    SCK(cudaMalloc(&d_lp, sizeof(double) * W));
    SCK(cudaMalloc(&d_prop, sizeof(double) * half * ndim));
    SCK(cudaMalloc(&d_lpnew, sizeof(double) * half));
    SCK(cudaMalloc(&d_lnz, sizeof(double) * half));
    SCK(cudaMalloc(&d_lnu, sizeof(double) * half));
    SCK(cudaMalloc(&d_lpnext, sizeof(double) * W));
    SCK(cudaMalloc(&d_nacc, sizeof(long long) * W));
    SCK(cudaMalloc(&d_step, sizeof(long long)));
    if (chain && nsteps > 0) SCK(cudaMalloc(&d_chain, sizeof(double) * (size_t)nsteps * W * ndim));
    if (lnprob && nsteps > 0) SCK(cudaMalloc(&d_lnp, sizeof(double) * (size_t)nsteps * W));
    SCK(cudaMemcpyAsync(d_pos, pos0, sizeof(double) * W * ndim, cudaMemcpyHostToDevice, st));
    SCK(cudaMemsetAsync(d_nacc, 0, sizeof(long long) * W, st));
    const long long s0 = step0;
    SCK(cudaMemcpyAsync(d_step, &s0, sizeof(long long), cudaMemcpyHostToDevice, st));
    // log-posterior of the starting ensemble, and one un-captured half-ensemble call so that every scratch buffer
    // the captured pipeline needs already has its final size
    // source-sharded run over several GPUs: every rank draws the same Philox proposals and the per-walker partials are
    // summed by the peer-memory kernel inside the captured update (identical bits on every rank keep the chains equal)
    const bool exchange = c->peer_connected && c->peer.world > 1;
    if (exchange && W > c->peer.wcap) { fail("lf_sampler_run: more walkers than the peer buffers hold"); return cleanup(); }
    // walker sharding (small catalogues, every rank holds ALL sources): a rank evaluates only its slice of the walkers and
    // contributes exact zeros for the others, so the same rank-ordered sum over ranks IS the all-gather of the slices
    // (x + 0 = x bit for bit, -inf + 0 = -inf) and the exchange kernel, its flags and the captured graph stay as they are
    const bool wshard = exchange && c->walker_shard;
    auto lnprob_all_ranks = [&](const double* th, long long nw, double* out) -> int {
        if (wshard) {
            const long long lo = nw * c->peer.rank / c->peer.world, hi = nw * (c->peer.rank + 1) / c->peer.world;
            if (cudaMemsetAsync(out, 0, sizeof(double) * nw, st) != cudaSuccess) return fail("lf_sampler_run: cudaMemsetAsync failed");
            if (hi > lo && launch_pipeline(c, th + lo * ndim, hi - lo, out + lo, st)) return 1;
        } else if (launch_pipeline(c, th, nw, out, st)) return 1;
        if (exchange) {
            if (peer_allreduce_launch(c, out, nw, st)) return 1;
        }
        return 0;
    };
    // the update with index `step` reads the log-posteriors from lp (even step) or lp_next (odd step) and writes the other
    if (lnprob_all_ranks(d_pos, W, (step0 & 1) ? d_lpnext : d_lp)) return cleanup();
    if (lnprob_all_ranks(d_pos, half, d_lpnew)) return cleanup();
    SCK(cudaStreamSynchronize(st));
    SamplerArgs sa;
    sa.W = (int)W; sa.half = half; sa.ndim = ndim; sa.a = a;
    sa.k0 = (uint32_t)seed; sa.k1 = (uint32_t)(seed >> 32);
    sa.step = d_step; sa.step0 = step0;
    sa.pos = d_pos; sa.lp = d_lp; sa.prop = d_prop; sa.lpnew = d_lpnew; sa.lnz = d_lnz; sa.lnu = d_lnu;
    sa.chain = d_chain; sa.lnp = d_lnp; sa.nacc = d_nacc; sa.lp_next = d_lpnext;
    const int T = 128, G = (half * ndim + T - 1) / T;
    const long long launches0 = c->launches;
    if (nsteps > 0) {
        SCK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        bool ok = true;
        for (int h = 0; h < 2 && ok; ++h) {
            k_stretch_propose<<<G, T, 0, st>>>(sa, h);
            ok = lnprob_all_ranks(d_prop, half, d_lpnew) == 0;
            k_stretch_accept<<<G, T, 0, st>>>(sa, h);
            c->launches += 2;
        }
        k_step_advance<<<1, 1, 0, st>>>(d_step);
        c->launches += 1;
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (!ok) return cleanup();
        SCK(ce);
        SCK(cudaGraphInstantiate(&exec, graph, 0));
        const long long per_step = c->launches - launches0;
        SCK(cudaEventRecord(c->ev0, st));
        for (int64_t t = 0; t < nsteps; ++t) {
            // a peer time-out is sticky and poisons every later exchange with NaN (all proposals rejected): stop
            // enqueueing as soon as the host-mapped flag shows it instead of replaying the update to the end
            if (exchange && c->peer_timeout_h && *(volatile int*)c->peer_timeout_h) {
                cudaStreamSynchronize(st);
                fail("lf_sampler_run: the peer-memory exchange timed out waiting for another rank at update " + std::to_string((long long)t));
                return cleanup();
            }
            SCK(cudaGraphLaunch(exec, st));
        }
        SCK(cudaEventRecord(c->ev1, st));
        c->launches = launches0 + per_step * nsteps;
    }
    if (chain && nsteps > 0) SCK(cudaMemcpyAsync(chain, d_chain, sizeof(double) * (size_t)nsteps * W * ndim, cudaMemcpyDeviceToHost, st));
    if (lnprob && nsteps > 0) SCK(cudaMemcpyAsync(lnprob, d_lnp, sizeof(double) * (size_t)nsteps * W, cudaMemcpyDeviceToHost, st));
    if (naccepted) SCK(cudaMemcpyAsync(naccepted, d_nacc, sizeof(long long) * W, cudaMemcpyDeviceToHost, st));
    if (pos_out) SCK(cudaMemcpyAsync(pos_out, d_pos, sizeof(double) * W * ndim, cudaMemcpyDeviceToHost, st));
    if (lnprob_out) SCK(cudaMemcpyAsync(lnprob_out, ((step0 + nsteps) & 1) ? d_lpnext : d_lp, sizeof(double) * W, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    SCK(cudaGetLastError());
    if (exchange && c->peer_timeout_h && *(volatile int*)c->peer_timeout_h) {
        fail("lf_sampler_run: the peer-memory exchange timed out waiting for another rank; the chain is invalid from that update on");
        return cleanup();
    }
    if (nsteps > 0) { float ms = 0.f; SCK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->sampler_ms = ms; }
#undef SCK
    rc = 0;
    return cleanup();
}

extern "C" int lf_set_walker_sharding(lf_ctx* c, int32_t enabled) {
    if (!c) return fail("lf_set_walker_sharding: null context");
    if (enabled && (c->ka.nshare != 1)) return fail("lf_set_walker_sharding: a walker-sharded context integrates every walker's quadrature (quadrature share must be (0, 1))");
    c->walker_shard = enabled != 0;
    return 0;
}

extern "C" int lf_sampler_last_ms(lf_ctx* c, double* ms) {
    if (!c || !ms) return fail("lf_sampler_last_ms: null argument");
    *ms = c->sampler_ms;
    return 0;
}

