// lf_peer.cu -- the multi-GPU exchange as one kernel over peer memory (CUDA IPC, P2P stores over NVLink, flags).
#include "lf_internal.cuh"

// ------------------------------------------------------------------------------------------------
// peer-memory all-reduce of the per-walker partials (one process per GPU, NVLink / NVSwitch)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(PEER_CHUNK) k_allreduce_p2p(PeerArgs a, double* __restrict__ vec, long long W) {
    const unsigned seq = *a.seq;
    const int par = (int)(seq & 1u);
    const int chunk = blockIdx.x;
    const long long w = (long long)chunk * PEER_CHUNK + threadIdx.x;
    const size_t slot = ((size_t)par * a.world + a.rank) * (size_t)a.wcap;
    // 1. push my values into my slot of every rank's buffer (coalesced 8-byte stores over NVLink; own buffer included)
    if (w < W) {
        const double v = vec[w];
        for (int r = 0; r < a.world; ++r) a.data[r][slot + w] = v;
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag for this chunk on every rank
    if (threadIdx.x < a.world)
        st_release_sys(a.flags[threadIdx.x] + ((size_t)par * a.world + a.rank) * a.nchunk_cap + chunk, seq);
    // 3. wait for every sender's flag on my own buffer (bounded spin: a dead peer must not hang the GPU).  A wait that
    //    expires -- or a time-out recorded by an EARLIER exchange (the flag is sticky) -- poisons this chunk's output
    //    with NaN: a stale partial sum can never be mistaken for a log-posterior (the samplers reject NaN), and the
    //    host raises on its next look at the flag.
    __shared__ int s_expired;
    if (threadIdx.x == 0) s_expired = *(volatile int*)a.timed_out;
    __syncthreads();
    if (threadIdx.x < a.world) {
        const unsigned* f = a.flags[a.rank] + ((size_t)par * a.world + threadIdx.x) * a.nchunk_cap + chunk;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(f) - seq) < 0) {
            if (clock64() - t0 > a.spin_limit) { atomicExch(a.timed_out, 1); s_expired = 1; break; }
            __nanosleep(100);
        }
    }
    __syncthreads();
    // 4. add the slots in rank order: the same sum, bit for bit, on every rank
    if (w < W) {
        double s = 0.0;
        for (int r = 0; r < a.world; ++r) s += a.data[a.rank][((size_t)par * a.world + r) * (size_t)a.wcap + w];
        vec[w] = s_expired ? __longlong_as_double(0x7ff8000000000000LL) : s;
    }
}
__global__ void k_seq_advance(unsigned* seq) { *seq += 1u; }

extern "C" int lf_peer_buffer_create(lf_ctx* c, int32_t rank, int32_t world, int64_t wcap, unsigned char handle_out[64]) {
    if (!c || !handle_out) return fail("lf_peer_buffer_create: null argument");
    if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world || wcap < 1) return fail("lf_peer_buffer_create: bad rank / world / capacity");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    CK(cudaSetDevice(c->device));
    if (c->peer_base) return fail("lf_peer_buffer_create: already created");
    const long long cap = (wcap + PEER_CHUNK - 1) / PEER_CHUNK * PEER_CHUNK;
    const int nchunk = (int)(cap / PEER_CHUNK);
    const size_t data_bytes = sizeof(double) * 2 * (size_t)world * cap;
    const size_t flag_bytes = sizeof(unsigned) * 2 * (size_t)world * nchunk;
    CK(cudaMalloc(&c->peer_base, data_bytes + flag_bytes));
    CK(cudaMemset(c->peer_base, 0, data_bytes + flag_bytes));
    CK(cudaMalloc(&c->peer_seq, sizeof(unsigned)));
    const unsigned one = 1u;
    CK(cudaMemcpy(c->peer_seq, &one, sizeof(unsigned), cudaMemcpyHostToDevice));
    // time-out flag in mapped pinned host memory: the kernel writes it (zero-copy) only when a wait expires, the host
    // reads it without a device round trip
    CK(cudaHostAlloc(&c->peer_timeout_h, sizeof(int), cudaHostAllocMapped));
    *c->peer_timeout_h = 0;
    CK(cudaHostGetDevicePointer(&c->peer_timeout, c->peer_timeout_h, 0));
    PeerArgs& p = c->peer;
    memset(&p, 0, sizeof(p));
    p.rank = rank; p.world = world; p.wcap = cap; p.nchunk_cap = nchunk;
    p.data[rank] = reinterpret_cast<double*>(c->peer_base);
    p.flags[rank] = reinterpret_cast<unsigned*>(c->peer_base + data_bytes);
    p.seq = c->peer_seq; p.timed_out = c->peer_timeout;
    p.spin_limit = PEER_SPIN_CLOCKS_DEFAULT;
    c->peer_data_bytes = data_bytes;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->peer_base));
    memcpy(handle_out, &h, 64);
    c->peer_connected = (world == 1);
    return 0;
}

extern "C" int lf_peer_buffer_connect(lf_ctx* c, const unsigned char* handles) {
    if (!c || !handles) return fail("lf_peer_buffer_connect: null argument");
    if (!c->peer_base) return fail("lf_peer_buffer_connect: call lf_peer_buffer_create first");
    CK(cudaSetDevice(c->device));
    PeerArgs& p = c->peer;
    for (int r = 0; r < p.world; ++r) {
        if (r == p.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * r, 64);
        bool absent = true;                          // an all-zero row: no peer is mapped for rank r (tests of the time-out
        for (int b = 0; b < 64; ++b) absent = absent && handles[64 * r + b] == 0;    // path); its stores fall into this rank's own buffer
        if (absent) { p.data[r] = p.data[p.rank]; p.flags[r] = p.flags[p.rank]; continue; }
        void* base = nullptr;
        CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_opened[r] = base;
        p.data[r] = reinterpret_cast<double*>(base);
        p.flags[r] = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(base) + c->peer_data_bytes);
    }
    c->peer_connected = true;
    return 0;
}

extern "C" int lf_allreduce_device(lf_ctx* c, double* d_vec, int64_t W, void* stream) {
    if (!c || (W > 0 && !d_vec)) return fail("lf_allreduce_device: null argument");
    if (!c->peer_base || !c->peer_connected) return fail("lf_allreduce_device: peer buffers are not connected");
    if (W > c->peer.wcap) return fail("lf_allreduce_device: vector longer than the peer buffer capacity");
    if (W <= 0) return 0;
    CK(cudaSetDevice(c->device));
    return peer_allreduce_launch(c, d_vec, W, (cudaStream_t)stream);
}

int peer_allreduce_launch(lf_ctx* c, double* d_vec, long long W, cudaStream_t st) {
    // sticky: once a wait has expired every later exchange is refused until the caller acknowledges it (lf_peer_status)
    if (c->peer_timeout_h && *(volatile int*)c->peer_timeout_h)
        return fail("peer-memory all-reduce: an earlier exchange timed out waiting for another rank; results since then are NaN");
    k_allreduce_p2p<<<(unsigned)((W + PEER_CHUNK - 1) / PEER_CHUNK), PEER_CHUNK, 0, st>>>(c->peer, d_vec, W);
    k_seq_advance<<<1, 1, 0, st>>>(c->peer_seq);
    c->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

void peer_release(lf_ctx* c) {
    for (int r = 0; r < PEER_MAX; ++r)
        if (c->peer_opened[r]) cudaIpcCloseMemHandle(c->peer_opened[r]);
    dfree(c->peer_base);
    dfree(c->peer_seq);
    if (c->peer_timeout_h) cudaFreeHost(c->peer_timeout_h);
    c->peer_timeout_h = nullptr;
}

extern "C" int lf_peer_status(lf_ctx* c, int32_t* timed_out) {
    if (!c || !timed_out) return fail("lf_peer_status: null argument");
    *timed_out = 0;
    if (!c->peer_timeout_h) return 0;
    *timed_out = *(volatile int*)c->peer_timeout_h;        // sticky: stays set (and every later exchange is refused / NaN) until lf_peer_reset
    return 0;
}

extern "C" int lf_peer_reset(lf_ctx* c) {
    if (!c) return fail("lf_peer_reset: null context");
    if (c->peer_timeout_h) *c->peer_timeout_h = 0;
    return 0;
}

extern "C" int lf_peer_set_timeout(lf_ctx* c, double seconds) {
    if (!c || !(seconds > 0.0)) return fail("lf_peer_set_timeout: need a context and seconds > 0");
    int khz = 1965000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
    c->peer.spin_limit = (long long)(seconds * 1.0e3 * (double)khz);
    return 0;
}

