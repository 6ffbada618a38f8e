// comm.cu -- tensor-parallel all-reduce behind B200_OP_ALLREDUCE (one process per GPU, one b200_ctx per process).
//
// Replaces the multi-GPU matmul driver of the reference (ggml_cuda_op_mul_mat over a split buffer, ggml-cuda.cu:1363-1671:
// row slices gathered on the main GPU with peer memcpy + an event per matmul, no CUDA graphs) with a Megatron-style split:
// wq/wk/wv and gate/up by output rows, wo and down by K, and ONE f32 sum over ranks of the [E, n_tok] partial results after
// wo and after down.  Two implementations behind the same op:
//   * one-shot peer-memory kernel (decode-sized vectors, <= B200_ONESHOT_MAX_BYTES): every rank stores its partial vector
//     into a slot of EVERY peer's exchange buffer over NVLink (plain st.global on cudaIpc-mapped peer pointers), publishes
//     an epoch flag with a system-scope release, spins on its own flags with acquire loads, then adds the world slots in
//     FIXED rank order (+ the optional residual): every rank computes bit-identical sums, there is no host round trip and
//     the kernel is CUDA-graph capturable (the epoch lives in device memory).  Two slot sets alternate by epoch parity, so
//     a rank that runs one all-reduce ahead never overwrites data a slower peer is still reading;
//   * NCCL ncclAllReduce(float, sum) for prefill-sized tensors, on the context's stream (graph capturable).  NCCL is
//     loaded at run time (dlopen libnccl.so.2: the one already in the process when torch is, the system one otherwise).
#include "common.cuh"
#include <dlfcn.h>

namespace {

constexpr size_t ONESHOT_MAX_BYTES = 256 * 1024;       // per-rank vector size served by the peer-memory kernel
constexpr int    ONESHOT_CTAS      = 16;
constexpr int    MAX_WORLD         = 8;

// ---- NCCL through dlopen (no link-time dependency; the C ABI stays free of NCCL types) ----
struct NcclId { char internal[128]; };
typedef int (*nccl_get_unique_id_t)(NcclId *);
typedef int (*nccl_comm_init_rank_t)(void **, int, NcclId, int);
typedef int (*nccl_all_reduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_comm_destroy_t)(void *);
typedef const char *(*nccl_get_error_string_t)(int);
struct NcclApi {
    void *h = nullptr;
    nccl_get_unique_id_t get_unique_id = nullptr;
    nccl_comm_init_rank_t comm_init_rank = nullptr;
    nccl_all_reduce_t all_reduce = nullptr;
    nccl_comm_destroy_t comm_destroy = nullptr;
    nccl_get_error_string_t err = nullptr;
} g_nccl;

bool nccl_load() {
    if (g_nccl.h) return true;
    const char *cands[] = {getenv("GGML_B200_NCCL_PATH"), "libnccl.so.2", "libnccl.so"};
    for (const char *c : cands) {
        if (!c) continue;
        void *h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (!h) continue;
        g_nccl.get_unique_id = (nccl_get_unique_id_t)dlsym(h, "ncclGetUniqueId");
        g_nccl.comm_init_rank = (nccl_comm_init_rank_t)dlsym(h, "ncclCommInitRank");
        g_nccl.all_reduce = (nccl_all_reduce_t)dlsym(h, "ncclAllReduce");
        g_nccl.comm_destroy = (nccl_comm_destroy_t)dlsym(h, "ncclCommDestroy");
        g_nccl.err = (nccl_get_error_string_t)dlsym(h, "ncclGetErrorString");
        if (g_nccl.get_unique_id && g_nccl.comm_init_rank && g_nccl.all_reduce && g_nccl.comm_destroy) { g_nccl.h = h; return true; }
        dlclose(h);
    }
    b200_set_error("comm: libnccl.so.2 not found (set GGML_B200_NCCL_PATH)");
    return false;
}

}  // namespace

// exchange buffer of one rank (device memory, cudaIpc-shared):
//   [ epoch : u32, pad to 128 B ][ flags : 2 sets x MAX_WORLD ranks x ONESHOT_CTAS u32 ][ slots : 2 sets x MAX_WORLD x ONESHOT_MAX_BYTES ]
struct b200_comm {
    int rank = 0, world = 1;
    void *nccl = nullptr;
    uint8_t *local = nullptr;                  // this rank's exchange buffer
    uint8_t *peer[MAX_WORLD] = {nullptr};      // mapped exchange buffers of all ranks (peer[rank] == local)
    bool peers_attached = false;
    int oneshot = 1;
    uint32_t *err_host = nullptr, *err_dev = nullptr;     // mapped pinned word: a kernel that gave up waiting for a peer sets it (comm_check)
};

namespace {

constexpr size_t OFF_FLAGS = 128;
constexpr size_t FLAGS_BYTES = 2 * MAX_WORLD * ONESHOT_CTAS * 4;
constexpr size_t OFF_SLOTS = (OFF_FLAGS + FLAGS_BYTES + 127) & ~(size_t)127;
constexpr size_t LL_MAX_BYTES = 64 * 1024;                   // payload per rank served by the flag-in-data (LL) kernel
constexpr size_t OFF_LL = OFF_SLOTS + 2 * MAX_WORLD * ONESHOT_MAX_BYTES;
constexpr size_t EXCH_BYTES = OFF_LL + 2 * MAX_WORLD * 2 * LL_MAX_BYTES;

struct OneShotParams {
    uint8_t *peer[MAX_WORLD];
    const float *src; const float *residual; float *dst;
    int n, rank, world;                        // n floats (multiple of 4)
    int use_pdl;
    uint32_t *err;                             // host-visible error word (peer timeout)
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t *p) { uint32_t v; asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// grid = ONESHOT_CTAS; CTA c owns floats [c*per, (c+1)*per) of the vector on every rank
__global__ void __launch_bounds__(256) b200_allreduce_oneshot_kernel(const OneShotParams p) {
    if (p.use_pdl) { pdl_trigger(); pdl_wait(); }
    uint8_t *local = p.peer[p.rank];
    const uint32_t epoch = *(volatile uint32_t *)local + 1;       // bumped by the last CTA at the end; same value on all ranks
    const int set = epoch & 1;
    const int n4 = p.n >> 2;
    const int per = (n4 + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(n4, lo + per);
    const float4 *src = (const float4 *)p.src;
    // 1. push my slice into slot [set][rank] of every rank (mine included)
    for (int r = 0; r < p.world; r++) {
        float4 *slot = (float4 *)(p.peer[r] + OFF_SLOTS + ((size_t)set * MAX_WORLD + p.rank) * ONESHOT_MAX_BYTES);
        for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) slot[i] = src[i];
    }
    __syncthreads();
    // 2. publish: flag [set][rank][cta] of every rank = epoch.  One system-scope fence per publishing thread AFTER the CTA barrier
    //    (cumulative over the whole CTA's stores, the cooperative-groups grid-sync pattern) instead of one per thread
    if (threadIdx.x < p.world) {
        uint32_t *f = (uint32_t *)(p.peer[threadIdx.x] + OFF_FLAGS) + ((size_t)set * MAX_WORLD + p.rank) * ONESHOT_CTAS + blockIdx.x;
        st_release_sys(f, epoch);                 // the release carries the (one) system-scope fence
    }
    // 3. wait for the same slice of every rank (bounded spin: a lost peer must not hang the GPU)
    if (threadIdx.x < p.world) {
        const uint32_t *f = (const uint32_t *)(local + OFF_FLAGS) + ((size_t)set * MAX_WORLD + threadIdx.x) * ONESHOT_CTAS + blockIdx.x;
        const long long t0 = clock64();
        while ((int)(ld_relaxed_sys(f) - epoch) < 0) { if (clock64() - t0 > (1ll << 32)) { if (p.err) *(volatile uint32_t *)p.err = 1u + threadIdx.x; break; } }     // plain polls ...
        (void)ld_acquire_sys(f);                                                                      // ... one acquire once the flag is there
    }
    __syncthreads();
    // 4. sum in fixed rank order (+ residual)
    const float4 *res = (const float4 *)p.residual;
    float4 *dst = (float4 *)p.dst;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        float4 a = __ldcg((const float4 *)(local + OFF_SLOTS + ((size_t)set * MAX_WORLD + 0) * ONESHOT_MAX_BYTES) + i);      // L2: peers wrote these lines
        for (int r = 1; r < p.world; r++) {
            const float4 b = __ldcg((const float4 *)(local + OFF_SLOTS + ((size_t)set * MAX_WORLD + r) * ONESHOT_MAX_BYTES) + i);
            a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
        }
        if (res) { const float4 b = res[i]; a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w); }
        dst[i] = a;
    }
    // 5. the last CTA to finish bumps the epoch (device-resident, so a captured graph replays correctly)
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t *done = (uint32_t *)(local + 4);
        __threadfence();
        if (atomicAdd(done, 1u) == gridDim.x - 1) { *done = 0; __threadfence(); *(volatile uint32_t *)local = epoch; }
    }
}

// Low-latency variant for decode-sized vectors (<= 64 KB): every 8-byte store carries 4 bytes of payload and the 4-byte epoch
// (the "LL" idea of NCCL's small-message protocol), so a receiver that sees the epoch in a word also sees its payload:
// no fence, no separate flag, no CTA barrier -- one NVLink traversal plus rank skew.  Slots: [2 sets][MAX_WORLD][2 x payload].
__device__ __forceinline__ void st_v4_sys(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_v4_sys(const void *p) {
    uint4 v; asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v;
}
__global__ void __launch_bounds__(256) b200_allreduce_ll_kernel(const OneShotParams p) {
    if (p.use_pdl) { pdl_trigger(); pdl_wait(); }
    uint8_t *local = p.peer[p.rank];
    const uint32_t epoch = *(volatile uint32_t *)local + 1;
    const int set = epoch & 1;
    const int n4 = p.n >> 2;
    const float4 *src = (const float4 *)p.src;
    const float4 *res = (const float4 *)p.residual;
    float4 *dst = (float4 *)p.dst;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const float4 v = src[i];
        const uint32_t x = __float_as_uint(v.x), y = __float_as_uint(v.y), z = __float_as_uint(v.z), w = __float_as_uint(v.w);
        for (int r = 0; r < p.world; r++) {
            uint8_t *slot = p.peer[r] + OFF_LL + ((size_t)set * MAX_WORLD + p.rank) * 2 * LL_MAX_BYTES + (size_t)i * 32;
            st_v4_sys(slot, x, epoch, y, epoch);
            st_v4_sys(slot + 16, z, epoch, w, epoch);
        }
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const long long t0 = clock64();
        for (int r = 0; r < p.world; r++) {                      // fixed rank order => bit-identical sums on every rank
            const uint8_t *slot = local + OFF_LL + ((size_t)set * MAX_WORLD + r) * 2 * LL_MAX_BYTES + (size_t)i * 32;
            uint4 lo, hi;
            do {
                lo = ld_v4_sys(slot); hi = ld_v4_sys(slot + 16);
                if (clock64() - t0 > (1ll << 32)) { if (p.err) *(volatile uint32_t *)p.err = 1u + r; break; }           // a lost peer must not hang the GPU
            } while (lo.y != epoch || lo.w != epoch || hi.y != epoch || hi.w != epoch);
            const float bx = __uint_as_float(lo.x), by = __uint_as_float(lo.z), bz = __uint_as_float(hi.x), bw = __uint_as_float(hi.z);
            if (r == 0) a = make_float4(bx, by, bz, bw);
            else { a.x = __fadd_rn(a.x, bx); a.y = __fadd_rn(a.y, by); a.z = __fadd_rn(a.z, bz); a.w = __fadd_rn(a.w, bw); }
        }
        if (res) { const float4 b = res[i]; a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w); }
        dst[i] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t *done = (uint32_t *)(local + 4);
        __threadfence();
        if (atomicAdd(done, 1u) == gridDim.x - 1) { *done = 0; __threadfence(); *(volatile uint32_t *)local = epoch; }
    }
}

__global__ void b200_add_f32_kernel(const float *a, const float *b, float *d, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = __fadd_rn(a[i], b[i]);
}

}  // namespace

extern "C" {

int b200_comm_unique_id(void *id128) {
    if (!id128 || !nccl_load()) return B200_ERR_FAILED;
    NcclId id;
    const int rc = g_nccl.get_unique_id(&id);
    if (rc) { b200_set_error("ncclGetUniqueId: %s", g_nccl.err ? g_nccl.err(rc) : "?"); return B200_ERR_FAILED; }
    memcpy(id128, &id, sizeof(id));
    return B200_OK;
}

int b200_comm_init(b200_ctx *ctx, const void *id128, int rank, int world) {
    if (!ctx || rank < 0 || world < 1 || rank >= world || world > MAX_WORLD) { b200_set_error("comm_init: rank %d world %d", rank, world); return B200_ERR_FAILED; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (ctx->comm) return B200_OK;
    b200_comm *c = new b200_comm();
    c->rank = rank; c->world = world;
    if (const char *e = getenv("GGML_B200_ONESHOT")) c->oneshot = atoi(e);
    if (world > 1) {
        if (!id128 || !nccl_load()) { delete c; return B200_ERR_FAILED; }
        NcclId id; memcpy(&id, id128, sizeof(id));
        const int rc = g_nccl.comm_init_rank(&c->nccl, world, id, rank);
        if (rc) { b200_set_error("ncclCommInitRank: %s", g_nccl.err ? g_nccl.err(rc) : "?"); delete c; return B200_ERR_FAILED; }
    }
    void *buf = nullptr;
    if (cudaMalloc(&buf, EXCH_BYTES) != cudaSuccess) { cudaGetLastError(); delete c; b200_set_error("comm_init: exchange buffer"); return B200_ERR_ALLOC; }
    CUDA_TRY(cudaMemset(buf, 0, OFF_SLOTS));
    CUDA_TRY(cudaMemset((uint8_t *)buf + OFF_LL, 0, EXCH_BYTES - OFF_LL));      // LL words must not hold a stale epoch
    c->local = (uint8_t *)buf;
    if (cudaHostAlloc((void **)&c->err_host, 64, cudaHostAllocMapped) == cudaSuccess && cudaHostGetDevicePointer((void **)&c->err_dev, c->err_host, 0) == cudaSuccess) *c->err_host = 0;
    else { cudaGetLastError(); c->err_host = c->err_dev = nullptr; }
    c->peer[rank] = c->local;
    ctx->comm = c;
    return B200_OK;
}

int b200_comm_rank(const b200_ctx *ctx) { return ctx && ctx->comm ? ctx->comm->rank : 0; }
int b200_comm_world(const b200_ctx *ctx) { return ctx && ctx->comm ? ctx->comm->world : 1; }

/* 64-byte cudaIpcMemHandle of this rank's exchange buffer; the host (torch.distributed, MPI, a socket) gathers them */
int b200_comm_peer_handle(b200_ctx *ctx, void *handle64) {
    if (!ctx || !ctx->comm || !handle64) return B200_ERR_FAILED;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, ctx->comm->local));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t");
    memcpy(handle64, &h, 64);
    return B200_OK;
}

/* handles: world x 64 bytes in rank order.  After this the one-shot kernel serves small all-reduces. */
int b200_comm_peer_attach(b200_ctx *ctx, const void *handles) {
    if (!ctx || !ctx->comm || !handles) return B200_ERR_FAILED;
    b200_comm *c = ctx->comm;
    CUDA_TRY(cudaSetDevice(ctx->device));
    for (int r = 0; r < c->world; r++) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)r * 64, 64);
        void *p = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[r] = (uint8_t *)p;
    }
    c->peers_attached = true;
    return B200_OK;
}

int b200_comm_destroy(b200_ctx *ctx) {
    if (!ctx || !ctx->comm) return B200_OK;
    b200_comm *c = ctx->comm;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < c->world; r++) if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->nccl && g_nccl.comm_destroy) g_nccl.comm_destroy(c->nccl);
    if (c->local) cudaFree(c->local);
    if (c->err_host) cudaFreeHost(c->err_host);
    delete c;
    ctx->comm = nullptr;
    return B200_OK;
}

}  // extern "C"

// called after a stream synchronisation: did an all-reduce give up waiting for a peer?  (the sums it produced are then garbage)
int comm_check(b200_ctx *ctx) {
    b200_comm *c = ctx ? ctx->comm : nullptr;
    if (!c || !c->err_host || *(volatile uint32_t *)c->err_host == 0) return B200_OK;
    const uint32_t who = *(volatile uint32_t *)c->err_host - 1;
    *(volatile uint32_t *)c->err_host = 0;
    b200_set_error("all-reduce: rank %d timed out waiting for rank %u (peer lost or not running the same step)", c->rank, who);
    return B200_ERR_FAILED;
}

bool supports_allreduce(const b200_op *op) {
    const b200_tensor &s = op->src[0], &d = op->dst;
    if (s.type != B200_TYPE_F32 || d.type != B200_TYPE_F32 || !tensor_is_contiguous(s) || !tensor_is_contiguous(d)) return false;
    if (tensor_nelements(s) != tensor_nelements(d)) return false;
    if (op->n_src > 1 && op->src[1].data) {
        const b200_tensor &r = op->src[1];
        if (r.type != B200_TYPE_F32 || !tensor_is_contiguous(r) || tensor_nelements(r) != tensor_nelements(s)) return false;
    }
    return true;
}

// dst = sum over ranks of src0 (+ src1, the residual, added once after the sum)
int op_allreduce(b200_ctx *ctx, const b200_op *op) {
    const int64_t n = tensor_nelements(op->src[0]);
    const float *src = (const float *)op->src[0].data;
    const float *res = op->n_src > 1 ? (const float *)op->src[1].data : nullptr;
    float *dst = (float *)op->dst.data;
    if (n == 0) return B200_OK;
    b200_comm *c = ctx->comm;
    const int world = c ? c->world : 1;
    const bool aligned = !(n & 3) && !((uintptr_t)src & 15) && !((uintptr_t)dst & 15) && !((uintptr_t)res & 15);
    if (world > 1 && c->oneshot && c->peers_attached && aligned && (size_t)n * 4 <= ONESHOT_MAX_BYTES) {
        OneShotParams p = {};
        for (int r = 0; r < world; r++) p.peer[r] = c->peer[r];
        p.src = src; p.residual = res; p.dst = dst; p.n = (int)n; p.rank = c->rank; p.world = world; p.use_pdl = ctx->opt_pdl; p.err = c->err_dev;
        // measured on Llama-3-70B decode (profiles/r1_scale.md): flag-in-data wins at 2 ranks (6.3 vs 8.9 us per all-reduce) and loses
        // at 8 (22 vs 15.9 us: twice the NVLink bytes and world x polling per element), so it is the default for 2 ranks only
        static const int use_ll = getenv("GGML_B200_ALLREDUCE_LL") ? atoi(getenv("GGML_B200_ALLREDUCE_LL")) : -1;
        const bool ll = (use_ll < 0 ? world <= 2 : use_ll != 0) && (size_t)n * 4 <= LL_MAX_BYTES;
        int ctas = (int)((n / 4 + 255) / 256);
        if (ctas > ONESHOT_CTAS) ctas = ONESHOT_CTAS;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)ctas); cfg.blockDim = dim3(256); cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = p.use_pdl ? 1 : 0;
        if (ll) CUDA_TRY(cudaLaunchKernelEx(&cfg, b200_allreduce_ll_kernel, p));
        else CUDA_TRY(cudaLaunchKernelEx(&cfg, b200_allreduce_oneshot_kernel, p));
        ctx->launches++;
        return B200_OK;
    }
    if (world > 1) {
        if (!c->nccl) { b200_set_error("allreduce: communicator not initialised"); return B200_ERR_FAILED; }
        const int rc = g_nccl.all_reduce(src, dst, (size_t)n, /*ncclFloat*/ 7, /*ncclSum*/ 0, c->nccl, ctx->stream);
        if (rc) { b200_set_error("ncclAllReduce: %s", g_nccl.err ? g_nccl.err(rc) : "?"); return B200_ERR_FAILED; }
        ctx->launches++;
        src = dst;
    } else if (!res) {
        if (src != dst) CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        return B200_OK;
    }
    if (res) {
        b200_add_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(src, res, dst, (int)n);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return B200_OK;
}
