// gemv.cu -- decode GEMV (batch <= 8) over GGUF-layout quantised weights.
//
// Replaces mul_mat_vec_q (ggml-cuda/mmvq.cu:130-204, one row per 128-thread CTA, 2/4-byte loads, q8_1
// activations) with a B200 design:
//   * persistent CTAs, one per SM; CTA c owns a contiguous range of rows, i.e. ONE contiguous byte range
//     of the weight tensor (rows are stored back to back in GGUF);
//   * a dedicated producer warp streams that byte range into a shared-memory ring with 1-D bulk async
//     copies (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP) -- the number of bytes in flight per SM
//     is ring size, independent of register pressure (Little: 6.5 TB/s x ~0.8 us => >= 36 KB per SM);
//   * consumer warps decode the blocks straight out of shared memory in their file layout
//     (gemv_items.cuh), int8 dp4a block dots against activations that were quantised exactly like the
//     CPU oracle does (quant.cu), warp-shuffle reduction, deterministic cross-warp combine;
//   * weights are requested with an L2 evict_first policy (streamed once per token), activations stay
//     L2 resident.
// Algorithmic bytes per launch: N*K*bpw (weights) + ncols*K*~1.07 (int8 activations) + ncols*N*4.
#include "common.cuh"
#include "gemv_items.cuh"

namespace {

using namespace gemv;

constexpr int NW = 8;                 // consumer warps per CTA
constexpr int MAX_STAGES = 16;

struct GemvParams {
    const uint8_t *W;                 // row 0 of the weight matrix
    uint32_t       rb;                // bytes per row
    int            N, K;
    const uint8_t *act;               // activation scratch (ActLayout), ncols columns
    ActLayout      L;
    float *        dst;
    size_t         dst_stride;        // elements between columns
    int            rs, wpr;           // rows per stage, warps per row (rs*wpr == NW)
    int            nstages;
    uint32_t       stage_bytes;
    int            w_const;           // weights are constant across launches (PDL may prefetch them early)
    int            use_pdl;
    // shared memory carve-up (bytes from the 128-aligned base)
    uint32_t off_cnt, off_part, off_aq, off_ad, off_as, off_ring;
    uint32_t aq_col, ad_col, as_col;  // per-column sizes in smem: bytes / floats / int16
    // debug (block sums)
    int32_t *dbgP, *dbgM;
};

template <int TYPE, int NC, bool DBG>
__global__ void __launch_bounds__((NW + 1) * 32, 1) gemv_kernel(const GemvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full  = (uint64_t *)smem;
    uint64_t *empty = full + MAX_STAGES;
    int *    cnt    = (int *)(smem + p.off_cnt);
    float *  part   = (float *)(smem + p.off_part);
    int8_t * s_aq   = (int8_t *)(smem + p.off_aq);
    float *  s_ad   = (float *)(smem + p.off_ad);
    int16_t *s_as   = (int16_t *)(smem + p.off_as);
    uint8_t *ring   = smem + p.off_ring;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, c = blockIdx.x;
    const int r0 = (int)((int64_t)p.N * c / G), r1 = (int)((int64_t)p.N * (c + 1) / G);
    const int nrows = r1 - r0;
    const int T = (nrows + p.rs - 1) / p.rs;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nstages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < MAX_STAGES * NW; i += blockDim.x) cnt[i] = 0;
    __syncthreads();

    if (warp == NW) {
        // ------------------------------------------------------------------ producer
        if (lane == 0 && T > 0) {
            if (p.use_pdl && !p.w_const) pdl_wait();
            const uint64_t pol = l2_policy_evict_first();
            for (int t = 0; t < T; t++) {
                const int s = t % p.nstages;
                if (t >= p.nstages) mbar_wait(&empty[s], ((t / p.nstages) - 1) & 1);
                const int row = r0 + t * p.rs;
                const int nr = min(p.rs, r1 - row);
                const uint8_t *src = p.W + (size_t)row * p.rb;
                const uint32_t extra = (uint32_t)((uintptr_t)src & 15);
                const uint32_t bytes = (extra + (uint32_t)nr * p.rb + 15u) & ~15u;
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_g2s_hint(ring + (size_t)s * p.stage_bytes, src - extra, bytes, &full[s], pol);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    if (T == 0) return;
    if (p.use_pdl) pdl_wait();
    constexpr int ITEM = Traits<TYPE>::ITEM;
    constexpr int ASTR = ITEM + 16;
    const int nitems = num_items<TYPE>(p.K);
    {   // activations: global scratch -> padded shared layout
        const int tid = threadIdx.x, nth = NW * 32;
        const int nvec = p.K / 16;                       // 16-byte vectors of int8 per column (K % 32 == 0)
        for (int col = 0; col < NC; col++) {
            const uint8_t *g = p.act + (size_t)col * p.L.col_bytes;
            for (int v = tid; v < nvec; v += nth) {
                const int e = v * 16;
                *(uint4 *)(s_aq + (size_t)col * p.aq_col + (size_t)(e / ITEM) * ASTR + (e % ITEM)) = *(const uint4 *)(g + e);
            }
            const int nd = p.K / (Traits<TYPE>::Q8K ? 256 : 32), ns = p.K / (Traits<TYPE>::Q8K ? 16 : 32);
            for (int i = tid; i < nd; i += nth) s_ad[(size_t)col * p.ad_col + i] = ((const float *)(g + p.L.off_d))[i];
            for (int i = tid; i < ns; i += nth) s_as[(size_t)col * p.as_col + i] = ((const int16_t *)(g + p.L.off_sums))[i];
        }
    }
    named_bar_sync(1, NW * 32);

    ActView A;
    A.q = s_aq; A.d = s_ad; A.s = s_as;
    A.q_stride = p.aq_col; A.d_stride = p.ad_col; A.s_stride = p.as_col;

    const int wrow = warp / p.wpr, wsub = warp % p.wpr;
    for (int t = 0; t < T; t++) {
        const int s = t % p.nstages;
        const int row = r0 + t * p.rs;
        const int nr = min(p.rs, r1 - row);
        mbar_wait(&full[s], (t / p.nstages) & 1);
        if (wrow < nr) {
            const uint32_t extra = (uint32_t)((uintptr_t)(p.W + (size_t)row * p.rb) & 15);
            const uint8_t *rowp = ring + (size_t)s * p.stage_bytes + extra + (size_t)wrow * p.rb;
            float acc[NC];
#pragma unroll
            for (int i = 0; i < NC; i++) acc[i] = 0.0f;
            DbgSink dbg;
            if (DBG) {
                const int nblk = p.K / Traits<TYPE>::BLOCK;
                dbg.P = p.dbgP + (size_t)(row + wrow) * nblk;
                dbg.M = p.dbgM + (size_t)(row + wrow) * nblk;
            }
            for (int it = wsub * 32 + lane; it < nitems; it += 32 * p.wpr) dot_item<TYPE, NC, DBG>(rowp, it, p.K, A, acc, dbg);
#pragma unroll
            for (int i = 0; i < NC; i++) acc[i] = warp_reduce_sum(acc[i]);
            if (p.wpr == 1) {
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < NC; i++) p.dst[(size_t)i * p.dst_stride + row + wrow] = acc[i];
                }
            } else {
                // deterministic cross-warp combine: last arriving warp sums the partials in warp order
                float *pp = part + ((size_t)(s * NW + wrow * p.wpr) * NC);
                int last = 0;
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < NC; i++) pp[wsub * NC + i] = acc[i];
                    __threadfence_block();
                    last = atomicAdd(&cnt[s * NW + wrow], 1) == p.wpr - 1;
                    if (last) {
                        __threadfence_block();
#pragma unroll
                        for (int i = 0; i < NC; i++) {
                            float v = 0.0f;
                            for (int w = 0; w < p.wpr; w++) v += ((volatile float *)pp)[w * NC + i];
                            p.dst[(size_t)i * p.dst_stride + row + wrow] = v;
                        }
                        cnt[s * NW + wrow] = 0;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (p.use_pdl) pdl_trigger();
}

template <int TYPE, int NC, bool DBG>
int launch_t(b200_ctx *ctx, const GemvParams &p, int grid, size_t smem_bytes) {
    auto kern = gemv_kernel<TYPE, NC, DBG>;
    static bool attr_set[8] = {false};   // per device
    if (!attr_set[ctx->device & 7]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        attr_set[ctx->device & 7] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((NW + 1) * 32);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    int nattr = 0;
    if (p.use_pdl) {
        attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[nattr].val.programmaticStreamSerializationAllowed = 1;
        nattr++;
    }
    cfg.attrs = attr;
    cfg.numAttrs = nattr;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
    ctx->launches++;
    return B200_OK;
}

template <int TYPE, bool DBG>
int launch_nc(b200_ctx *ctx, const GemvParams &p, int ncols, int grid, size_t smem_bytes) {
    switch (ncols) {
        case 1: return launch_t<TYPE, 1, DBG>(ctx, p, grid, smem_bytes);
        case 2: return launch_t<TYPE, 2, DBG>(ctx, p, grid, smem_bytes);
        case 3: return launch_t<TYPE, 3, DBG>(ctx, p, grid, smem_bytes);
        case 4: return launch_t<TYPE, 4, DBG>(ctx, p, grid, smem_bytes);
        default: b200_set_error("gemv: ncols=%d", ncols); return B200_ERR_UNSUPPORTED;
    }
}

int item_elems(int type) { return (type == B200_TYPE_Q4_K || type == B200_TYPE_Q5_K) ? 64 : 128; }

// choose stage geometry + smem carve-up for `ncols` columns; returns false if it does not fit
bool plan(const b200_ctx *ctx, int type, uint32_t rb, int64_t K, int ncols, GemvParams &p, size_t &smem_bytes) {
    const int q8k = b200_act_mode_q8k(type);
    const int item = item_elems(type);
    const int64_t nitems = (K + item - 1) / item;
    p.aq_col = (uint32_t)(nitems * (item + 16));
    p.ad_col = (uint32_t)(((K / (q8k ? 256 : 32)) + 3) & ~3);
    p.as_col = (uint32_t)(((K / (q8k ? 16 : 32)) + 7) & ~7);
    uint32_t off = 2 * MAX_STAGES * 8;
    p.off_cnt = off;  off += MAX_STAGES * NW * 4;
    p.off_part = off; off += MAX_STAGES * NW * 4 * 4;       // NC <= 4 floats per (stage, warp)
    p.off_aq = off;   off += p.aq_col * ncols;
    off = (off + 15) & ~15u;
    p.off_ad = off;   off += p.ad_col * 4 * ncols;
    p.off_as = off;   off += p.as_col * 2 * ncols;
    off = (off + 127) & ~127u;
    p.off_ring = off;
    const size_t budget = ctx->smem_optin;
    if (off + 2 * 1024 > budget) return false;
    const size_t ring_budget = budget - off;
    // rows per stage: largest power of two <= NW whose stage is <= ~32 KB and still leaves >= 3 stages
    int rs = NW;
    while (rs > 1 && ((size_t)rs * rb > 32 * 1024 || ((size_t)rs * rb + 144) * 3 > ring_budget)) rs >>= 1;
    const uint32_t stage_bytes = (uint32_t)(((size_t)rs * rb + 32 + 127) & ~(size_t)127);   // +16 misalignment, +16 over-read
    int ns = (int)(ring_budget / stage_bytes);
    if (ns < 2) return false;
    if (ns > MAX_STAGES) ns = MAX_STAGES;
    p.rs = rs; p.wpr = NW / rs; p.nstages = ns; p.stage_bytes = stage_bytes;
    smem_bytes = off + (size_t)ns * stage_bytes;
    return true;
}

template <bool DBG>
int launch_type(b200_ctx *ctx, int type, const GemvParams &p, int ncols, int grid, size_t smem) {
    switch (type) {
        case B200_TYPE_Q4_0: return launch_nc<T_Q4_0, DBG>(ctx, p, ncols, grid, smem);
        case B200_TYPE_Q8_0: return launch_nc<T_Q8_0, DBG>(ctx, p, ncols, grid, smem);
        case B200_TYPE_Q4_K: return launch_nc<T_Q4_K, DBG>(ctx, p, ncols, grid, smem);
        case B200_TYPE_Q5_K: return launch_nc<T_Q5_K, DBG>(ctx, p, ncols, grid, smem);
        case B200_TYPE_Q6_K: return launch_nc<T_Q6_K, DBG>(ctx, p, ncols, grid, smem);
        default: b200_set_error("gemv: type %d", type); return B200_ERR_UNSUPPORTED;
    }
}

}  // namespace

int gemv_max_cols(const b200_ctx *ctx, int type, size_t rb, int64_t K) {
    GemvParams p; size_t smem;
    for (int nc = 4; nc >= 1; nc--) if (plan(ctx, type, (uint32_t)rb, K, nc, p, smem)) return nc;
    return 0;
}

int launch_gemv(b200_ctx *ctx, int type, const uint8_t *W, size_t row_bytes, int64_t N, int64_t K, const uint8_t *act, int ncols,
                float *dst, size_t dst_col_stride, bool w_const) {
    if (N <= 0 || ncols <= 0) return B200_OK;
    const int q8k = b200_act_mode_q8k(type);
    const ActLayout L = ActLayout::make(q8k, K);
    if ((type == B200_TYPE_Q4_K || type == B200_TYPE_Q5_K) && (((uintptr_t)W & 15) || (row_bytes & 15))) {
        b200_set_error("gemv: K-quant rows must be 16-byte aligned");
        return B200_ERR_UNSUPPORTED;
    }
    const int maxc = gemv_max_cols(ctx, type, row_bytes, K);
    if (maxc == 0) { b200_set_error("gemv: K=%lld does not fit shared memory", (long long)K); return B200_ERR_UNSUPPORTED; }
    for (int c0 = 0; c0 < ncols; c0 += maxc) {
        const int nc = ncols - c0 < maxc ? ncols - c0 : maxc;
        GemvParams p = {};
        size_t smem = 0;
        plan(ctx, type, (uint32_t)row_bytes, K, nc, p, smem);
        p.W = W; p.rb = (uint32_t)row_bytes; p.N = (int)N; p.K = (int)K;
        p.act = act + (size_t)c0 * L.col_bytes; p.L = L;
        p.dst = dst + (size_t)c0 * dst_col_stride; p.dst_stride = dst_col_stride;
        p.w_const = w_const ? 1 : 0;
        p.use_pdl = ctx->opt_pdl;
        const int grid = (int)(N < ctx->sm_count ? N : ctx->sm_count);
        int rc = launch_type<false>(ctx, type, p, nc, grid, smem);
        if (rc) return rc;
    }
    return B200_OK;
}

// test hook: exact integer block sums computed by the SAME item decoders through the SAME pipeline
extern "C" int b200_block_sums(b200_ctx *ctx, int32_t type, const void *W, const float *x, int64_t N, int64_t K, int32_t *P, int32_t *M) {
    if (!ctx || !b200_type_is_quant(type)) { b200_set_error("block_sums: type %d", type); return B200_ERR_UNSUPPORTED; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int q8k = b200_act_mode_q8k(type);
    const ActLayout L = ActLayout::make(q8k, K);
    const int64_t nblk = K / b200_type_block_elems(type);
    uint8_t *scratch = (uint8_t *)ctx->get_scratch(SCRATCH_ACT, L.col_bytes + (size_t)N * 4);
    if (!scratch) return B200_ERR_ALLOC;
    float *dummy = (float *)(scratch + L.col_bytes);
    CUDA_TRY(cudaMemsetAsync(P, 0, (size_t)N * nblk * 4, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(M, 0, (size_t)N * nblk * 4, ctx->stream));
    int rc = launch_quantize_act(ctx, q8k, x, (size_t)K * 4, K, 1, scratch);
    if (rc) return rc;
    GemvParams p = {};
    size_t smem = 0;
    const size_t rb = b200_row_bytes(type, K);
    if (!plan(ctx, type, (uint32_t)rb, K, 1, p, smem)) { b200_set_error("block_sums: K too large"); return B200_ERR_UNSUPPORTED; }
    p.W = (const uint8_t *)W; p.rb = (uint32_t)rb; p.N = (int)N; p.K = (int)K;
    p.act = scratch; p.L = L; p.dst = dummy; p.dst_stride = (size_t)N;
    p.dbgP = P; p.dbgM = M;
    const int grid = (int)(N < ctx->sm_count ? N : ctx->sm_count);
    return launch_type<true>(ctx, type, p, 1, grid, smem);
}
