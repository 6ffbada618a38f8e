// gemv.cu -- decode GEMV (batch <= 8) over GGUF-layout quantised weights.
//
// Replaces mul_mat_vec_q (ggml-cuda/mmvq.cu:130-204: one row per 128-thread CTA, 2/4-byte loads, q8_1
// activations quantised by a separate kernel) with a B200 design:
//   * persistent CTAs, one per SM; CTA c owns a contiguous range of rows, i.e. ONE contiguous byte range
//     of the weight tensor (rows are stored back to back in GGUF);
//   * a dedicated producer warp streams that byte range into a deep shared-memory ring with 1-D bulk async
//     copies (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP), L2 evict_first -- bytes in flight per SM
//     = ring size (~190 KB), independent of register pressure (Little: 6.5 TB/s x ~1 us => >= 44 KB/SM);
//   * every ring stage (1-4 whole rows) is OWNED by one consumer warp: it waits on that stage's mbarrier,
//     decodes the blocks straight out of shared memory in their file layout (gemv_items.cuh) with int8
//     dp4a block dots, shuffle-reduces, stores, releases the stage.  No cross-warp synchronisation at all,
//     and the result of a row does not depend on scheduling (bit-reproducible);
//   * the activation quantisation (q8_K / q8_0, bit-exact vs the CPU oracle) is fused into the prologue:
//     each CTA quantises the f32 activations itself while the producer is already streaming weights
//     (optionally after an on-the-fly rms_norm*weight or silu(gate)*up), so a matmul is ONE launch.
// Algorithmic bytes per launch: N*K*bpw (weights) + ncols*K*4 (f32 activations) + ncols*N*4 (output).
#include "common.cuh"
#include "gemv_items.cuh"
#include "quant_warp.cuh"

namespace {

using namespace gemv;

constexpr int MAX_STAGES = 48;

enum { ACT_PREQ = 0, ACT_F32 = 1, ACT_F32_NORM = 2, ACT_F32_SWIGLU = 3 };

struct GemvParams {
    const uint8_t *W;                 // row 0 of the weight matrix
    uint32_t       rb;                // bytes per row
    int            N, K;
    int            act_mode;
    const uint8_t *act;               // ACT_PREQ: activation scratch (ActLayout), ncols columns
    ActLayout      L;
    const float *  x;                 // ACT_F32*: f32 activations, column stride x_stride bytes
    size_t         x_stride;
    const float *  x2;                // ACT_F32_NORM: norm weight [K]; ACT_F32_SWIGLU: `up` (same strides as x)
    float          eps;
    float *        dst;
    size_t         dst_stride;        // elements between columns
    const float *  residual;          // optional: dst = W.x + residual (same layout as dst)
    int            nstages;
    uint32_t       stage_bytes;
    int            w_const;           // weights are constant across launches (PDL may prefetch them early)
    int            use_pdl;
    // shared memory carve-up (bytes from the 128-aligned base)
    uint32_t off_aq, off_ad, off_as, off_ring;
    uint32_t aq_col, ad_col, as_col;  // per-column sizes in smem: bytes / floats / int16
    // debug (block sums)
    int32_t *dbgP, *dbgM;
    unsigned long long *prof;   // optional [grid][16] globaltimer stamps (tools/gemv_prof.py)
};

// ---- fused prologue: f32 activations -> quantised shared-memory layout --------------------------------
template <int TYPE>
__device__ __forceinline__ void quantize_col_to_smem(const GemvParams &p, int col, int8_t *s_aq, float *s_ad, int16_t *s_as,
                                                     int cwarp, int ncw, int lane, float norm_scale) {
    constexpr int ITEM = Traits<TYPE>::ITEM, ASTR = ITEM + 16;
    const float *xp = (const float *)((const char *)p.x + (size_t)col * p.x_stride);
    const float *up = p.act_mode == ACT_F32_SWIGLU ? (const float *)((const char *)p.x2 + (size_t)col * p.x_stride) : nullptr;
    const int nchunk = (p.K + 255) / 256;
    for (int b0 = cwarp; b0 < nchunk; b0 += 4 * ncw) {
        float4 ra[4][2], rb2[4][2];     // up to 4 chunks (x 2 operands) of loads in flight per warp
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e0 = (b0 + u * ncw) * 256 + lane * 8;
            if (b0 + u * ncw < nchunk && e0 < p.K) {
                ra[u][0] = *(const float4 *)(xp + e0); ra[u][1] = *(const float4 *)(xp + e0 + 4);
                if (p.act_mode == ACT_F32_NORM) { rb2[u][0] = *(const float4 *)(p.x2 + e0); rb2[u][1] = *(const float4 *)(p.x2 + e0 + 4); }
                else if (p.act_mode == ACT_F32_SWIGLU) { rb2[u][0] = *(const float4 *)(up + e0); rb2[u][1] = *(const float4 *)(up + e0 + 4); }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int b = b0 + u * ncw;
            if (b >= nchunk) break;
            const int e0 = b * 256 + lane * 8;
            const bool valid = e0 < p.K;
            float v[8];
            if (valid) {
                v[0] = ra[u][0].x; v[1] = ra[u][0].y; v[2] = ra[u][0].z; v[3] = ra[u][0].w;
                v[4] = ra[u][1].x; v[5] = ra[u][1].y; v[6] = ra[u][1].z; v[7] = ra[u][1].w;
                if (p.act_mode == ACT_F32_NORM || p.act_mode == ACT_F32_SWIGLU) {
                    const float w[8] = {rb2[u][0].x, rb2[u][0].y, rb2[u][0].z, rb2[u][0].w, rb2[u][1].x, rb2[u][1].y, rb2[u][1].z, rb2[u][1].w};
                    if (p.act_mode == ACT_F32_NORM) {
#pragma unroll
                        for (int j = 0; j < 8; j++) v[j] = __fmul_rn(__fmul_rn(v[j], norm_scale), w[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; j++) v[j] = __fmul_rn(__fdiv_rn(v[j], 1.0f + expf(-v[j])), w[j]);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = 0.0f;
            }
            uint2 qp;
            if (Traits<TYPE>::Q8K) {
                float d; int pair;
                warp_quant_q8k(v, lane, qp, d, pair);
                if ((lane & 1) == 0) s_as[(size_t)col * p.as_col + b * 16 + (lane >> 1)] = (int16_t)pair;
                if (lane == 0) s_ad[(size_t)col * p.ad_col + b] = d;
            } else {
                float d16; int bsum;
                warp_quant_q80(v, qp, d16, bsum);
                if (valid && (lane & 3) == 0) {
                    s_ad[(size_t)col * p.ad_col + (e0 >> 5)] = d16;
                    s_as[(size_t)col * p.as_col + (e0 >> 5)] = (int16_t)bsum;
                }
            }
            if (valid) *(uint2 *)(s_aq + (size_t)col * p.aq_col + (size_t)(e0 / ITEM) * ASTR + (e0 % ITEM)) = qp;
        }
    }
}

template <int TYPE, int NC, int RPW, bool DBG>
__global__ void __launch_bounds__(288, 2) gemv_kernel(const GemvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full  = (uint64_t *)smem;
    uint64_t *empty = full + MAX_STAGES;
    float *   s_red = (float *)(empty + MAX_STAGES);       // 32 floats: rms_norm reduction
    volatile int *s_seq = (volatile int *)(s_red + 32);    // chunk index currently held by each stage
    int8_t *  s_aq  = (int8_t *)(smem + p.off_aq);
    float *   s_ad  = (float *)(smem + p.off_ad);
    int16_t * s_as  = (int16_t *)(smem + p.off_as);
    uint8_t * ring  = smem + p.off_ring;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = (blockDim.x >> 5) - 1;                 // consumer warps; the last warp is the producer
#define PROF(slot) do { if (p.prof && lane == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.prof[(size_t)blockIdx.x * 16 + (slot)] = t_; } } while (0)
    if (warp == 0) PROF(0);
    const int G = gridDim.x, c = blockIdx.x;
    const int r0 = (int)((int64_t)p.N * c / G), r1 = (int)((int64_t)p.N * (c + 1) / G);
    const int nrows = r1 - r0;
    const int nchunks = (nrows + RPW - 1) / RPW;

    if (threadIdx.x < p.nstages) { mbar_init(&full[threadIdx.x], 1); mbar_init(&empty[threadIdx.x], 1); }
    mbar_fence_init();
    __syncthreads();
    if (p.use_pdl) pdl_trigger();     // dependents may start their own prologue / weight prefetch as SMs free up
    if (nchunks == 0) return;
    if (warp == 0) PROF(1);

    if (warp == ncw) {
        // ------------------------------------------------------------------ producer
        // lanes issue the copies of `rw` consecutive chunks in parallel (a single thread tops out near 250 ns per
        // copy, i.e. ~5 TB/s chip-wide); rw <= nstages so no lane waits on a stage another lane of the same round fills
        if (p.use_pdl && !p.w_const) pdl_wait();
        const uint64_t pol = l2_policy_evict_first();
        const int rw = min(32, p.nstages);
        for (int base = 0; base < nchunks; base += rw) {
            const int i = base + lane;
            if (lane < rw && i < nchunks) {
                const int s = i % p.nstages;
                if (i >= p.nstages) mbar_wait(&empty[s], ((i / p.nstages) - 1) & 1);
                const int row = r0 + i * RPW;
                const int nr = min(RPW, r1 - row);
                const uint8_t *src = p.W + (size_t)row * p.rb;
                const uint32_t extra = (uint32_t)((uintptr_t)src & 15);
                const uint32_t bytes = (extra + (uint32_t)nr * p.rb + 15u) & ~15u;
                s_seq[s] = i;                                   // published by the release of the arrive below
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_g2s_hint(ring + (size_t)s * p.stage_bytes, src - extra, bytes, &full[s], pol);
            }
            if (base == 0) PROF(2);
        }
        PROF(3);
        return;
    }

    // ---------------------------------------------------------------------- consumers: prologue
    if (p.use_pdl) pdl_wait();
    constexpr int ITEM = Traits<TYPE>::ITEM;
    constexpr int ASTR = ITEM + 16;
    const int nitems = num_items<TYPE>(p.K);
    const int nct = ncw * 32;
    if (p.act_mode == ACT_PREQ) {
        const int tid = threadIdx.x;
        const int nvec = p.K / 16;
        for (int col = 0; col < NC; col++) {
            const uint8_t *g = p.act + (size_t)col * p.L.col_bytes;
            for (int v = tid; v < nvec; v += nct) {
                const int e = v * 16;
                *(uint4 *)(s_aq + (size_t)col * p.aq_col + (size_t)(e / ITEM) * ASTR + (e % ITEM)) = *(const uint4 *)(g + e);
            }
            const int nd = p.K / (Traits<TYPE>::Q8K ? 256 : 32), ns = p.K / (Traits<TYPE>::Q8K ? 16 : 32);
            for (int i = tid; i < nd; i += nct) s_ad[(size_t)col * p.ad_col + i] = ((const float *)(g + p.L.off_d))[i];
            for (int i = tid; i < ns; i += nct) s_as[(size_t)col * p.as_col + i] = ((const int16_t *)(g + p.L.off_sums))[i];
        }
    } else {
        for (int col = 0; col < NC; col++) {
            float norm_scale = 1.0f;
            if (p.act_mode == ACT_F32_NORM) {
                // rms_norm exactly as glue.cu / the CPU oracle: sum of squares in double, fixed reduction order
                const float *xp = (const float *)((const char *)p.x + (size_t)col * p.x_stride);
                double s = 0.0;
                for (int i = threadIdx.x; i < p.K; i += nct) { const float v = xp[i]; s += (double)__fmul_rn(v, v); }
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                double *sd = (double *)s_red;
                named_bar_sync(1, nct);
                if (lane == 0) sd[warp] = s;
                named_bar_sync(1, nct);
                double t = 0.0;
                for (int i = 0; i < ncw; i++) t += sd[i];
                const float mean = (float)(t / (double)p.K);
                norm_scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, p.eps)));
            }
            quantize_col_to_smem<TYPE>(p, col, s_aq, s_ad, s_as, warp, ncw, lane, norm_scale);
        }
    }
    if (warp == 0) PROF(4);
    named_bar_sync(1, nct);
    if (warp == 0) PROF(5);

    ActView A;
    A.q = s_aq; A.d = s_ad; A.s = s_as;
    A.q_stride = p.aq_col; A.d_stride = p.ad_col; A.s_stride = p.as_col;

    // ---------------------------------------------------------------------- consumers: main loop
    for (int i = warp; i < nchunks; i += ncw) {
        const int s = i % p.nstages;
        const int row = r0 + i * RPW;
        const int nr = min(RPW, r1 - row);
        // A warp that runs far ahead of the owner of this slot's previous use could be fooled by the 1-bit phase
        // parity (use u and u+2 look alike); the stage's sequence number disambiguates.
        do { mbar_wait(&full[s], (i / p.nstages) & 1); } while (s_seq[s] != i);
        if (i == 0) PROF(6);
        const uint32_t extra = (uint32_t)((uintptr_t)(p.W + (size_t)row * p.rb) & 15);
        const uint8_t *rowp = ring + (size_t)s * p.stage_bytes + extra;
        float acc[RPW][NC];
#pragma unroll
        for (int r = 0; r < RPW; r++)
#pragma unroll
            for (int j = 0; j < NC; j++) acc[r][j] = 0.0f;
        DbgSink dbg[RPW];
        if (DBG) {
            const int nblk = p.K / Traits<TYPE>::BLOCK;
#pragma unroll
            for (int r = 0; r < RPW; r++) { dbg[r].P = p.dbgP + (size_t)(row + r) * nblk; dbg[r].M = p.dbgM + (size_t)(row + r) * nblk; }
        }
        constexpr bool TWO = (TYPE == T_Q4_K || TYPE == T_Q5_K) && RPW * NC <= 2;   // second item in flight only where registers allow
        for (int it = lane; it < nitems; it += TWO ? 64 : 32) {
            const bool two = TWO && it + 32 < nitems;
#pragma unroll
            for (int j = 0; j < NC; j++) {
                ActRegs<TYPE> ar0, ar1;
                load_act<TYPE>(A, j, it, ar0);
                if (two) load_act<TYPE>(A, j, it + 32, ar1);
#pragma unroll
                for (int r = 0; r < RPW; r++)
                    if (r < nr) {
                        const DbgSink ds = j == 0 ? dbg[r] : DbgSink{nullptr, nullptr};
                        float v = item_dot<TYPE, DBG>(rowp + (size_t)r * p.rb, it, p.K, ar0, ds);
                        if (two) v += item_dot<TYPE, DBG>(rowp + (size_t)r * p.rb, it + 32, p.K, ar1, ds);
                        acc[r][j] += v;
                    }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);       // stage bytes are in registers/accumulated: release before reducing
#pragma unroll
        for (int r = 0; r < RPW; r++)
#pragma unroll
            for (int j = 0; j < NC; j++) acc[r][j] = warp_reduce_sum(acc[r][j]);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < RPW; r++)
                if (r < nr) {
#pragma unroll
                    for (int j = 0; j < NC; j++) {
                        const size_t o = (size_t)j * p.dst_stride + row + r;
                        float v = acc[r][j];
                        if (p.residual) v = __fadd_rn(v, p.residual[o]);
                        p.dst[o] = v;
                    }
                }
        }
        if (i == 0) PROF(7);
        if (i + ncw >= nchunks && warp == (nchunks - 1) % ncw) PROF(8);
    }
#undef PROF
}

int item_elems(int type) { return (type == B200_TYPE_Q4_K || type == B200_TYPE_Q5_K) ? 64 : 128; }

// stage geometry + smem carve-up for `ncols` columns with `rpw` rows per stage; false if it does not fit
int g_gemv_warps = 0, g_gemv_stage_kb = 0, g_gemv_smem_kb = 0, g_gemv_ctas = 0;     // tuning overrides (env GGML_B200_GEMV_*)

bool plan(const b200_ctx *ctx, int type, uint32_t rb, int64_t K, int ncols, int rpw, int ctas_per_sm, GemvParams &p, size_t &smem_bytes) {
    const int q8k = b200_act_mode_q8k(type);
    const int item = item_elems(type);
    const int64_t nitems = (K + item - 1) / item;
    p.aq_col = (uint32_t)(nitems * (item + 16));
    p.ad_col = (uint32_t)(((K / (q8k ? 256 : 32)) + 4 + 3) & ~3);
    p.as_col = (uint32_t)(((K / (q8k ? 16 : 32)) + 8 + 7) & ~7);
    uint32_t off = 2 * MAX_STAGES * 8 + 128 + MAX_STAGES * 4 + 64;
    p.off_aq = off;   off += p.aq_col * ncols;
    off = (off + 15) & ~15u;
    p.off_ad = off;   off += p.ad_col * 4 * ncols;
    off = (off + 15) & ~15u;
    p.off_as = off;   off += p.as_col * 2 * ncols;
    off = (off + 127) & ~127u;
    p.off_ring = off;
    size_t budget = ctx->smem_optin;
    if ((size_t)off + 24 * 1024 <= (size_t)(g_gemv_smem_kb > 0 ? g_gemv_smem_kb : 112) * 1024) budget = (size_t)(g_gemv_smem_kb > 0 ? g_gemv_smem_kb : 112) * 1024;
    if ((size_t)off + 1024 > budget) return false;
    const size_t ring_budget = budget - off;
    const uint32_t stage_bytes = (uint32_t)(((size_t)rpw * rb + 32 + 127) & ~(size_t)127);   // +16 misalignment, +16 over-read
    int ns = (int)(ring_budget / stage_bytes);
    if (ns < 2) return false;
    if (ns > MAX_STAGES) ns = MAX_STAGES;
    p.nstages = ns; p.stage_bytes = stage_bytes;
    smem_bytes = off + (size_t)ns * stage_bytes;
    return true;
}

template <int TYPE, int NC, int RPW, bool DBG>
int launch_t(b200_ctx *ctx, const GemvParams &p, int grid, int nwarps, size_t smem_bytes) {
    auto kern = gemv_kernel<TYPE, NC, RPW, DBG>;
    static bool attr_set[16] = {false};   // per device
    if (!attr_set[ctx->device & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        attr_set[ctx->device & 15] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)(nwarps + 1) * 32);
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
int launch_shape(b200_ctx *ctx, const GemvParams &p, int ncols, int rpw, int grid, int nwarps, size_t smem) {
    if (ncols == 1) {
        if (rpw == 4) return launch_t<TYPE, 1, 4, DBG>(ctx, p, grid, nwarps, smem);
        if (rpw == 2) return launch_t<TYPE, 1, 2, DBG>(ctx, p, grid, nwarps, smem);
        return launch_t<TYPE, 1, 1, DBG>(ctx, p, grid, nwarps, smem);
    }
    if (DBG) return B200_ERR_UNSUPPORTED;
    if constexpr (!DBG) {
        switch (ncols) {
            case 2: return rpw >= 2 ? launch_t<TYPE, 2, 2, false>(ctx, p, grid, nwarps, smem) : launch_t<TYPE, 2, 1, false>(ctx, p, grid, nwarps, smem);
            case 3: return launch_t<TYPE, 3, 1, false>(ctx, p, grid, nwarps, smem);
            case 4: return launch_t<TYPE, 4, 1, false>(ctx, p, grid, nwarps, smem);
        }
    }
    b200_set_error("gemv: ncols=%d", ncols);
    return B200_ERR_UNSUPPORTED;
}

template <bool DBG>
int launch_type(b200_ctx *ctx, int type, const GemvParams &p, int ncols, int rpw, int grid, int nwarps, size_t smem) {
    switch (type) {
        case B200_TYPE_Q4_0: return launch_shape<T_Q4_0, DBG>(ctx, p, ncols, rpw, grid, nwarps, smem);
        case B200_TYPE_Q8_0: return launch_shape<T_Q8_0, DBG>(ctx, p, ncols, rpw, grid, nwarps, smem);
        case B200_TYPE_Q4_K: return launch_shape<T_Q4_K, DBG>(ctx, p, ncols, rpw, grid, nwarps, smem);
        case B200_TYPE_Q5_K: return launch_shape<T_Q5_K, DBG>(ctx, p, ncols, rpw, grid, nwarps, smem);
        case B200_TYPE_Q6_K: return launch_shape<T_Q6_K, DBG>(ctx, p, ncols, rpw, grid, nwarps, smem);
        default: b200_set_error("gemv: type %d", type); return B200_ERR_UNSUPPORTED;
    }
}


// rows per stage for a given column count: keep a stage around 8-12 KB so the ring holds >= 16 stages
int choose_rpw(uint32_t rb, int ncols, int64_t rows_per_cta) {
    const uint32_t target = (uint32_t)(g_gemv_stage_kb > 0 ? g_gemv_stage_kb : 10) * 1024;
    const int maxr = ncols == 1 ? 4 : (ncols == 2 ? 2 : 1);
    int r = maxr;
    // small stages when a CTA owns few rows (every consumer warp should get >= 2 chunks), bounded stage bytes otherwise
    while (r > 1 && ((size_t)r * rb > target || rows_per_cta / r < 16)) r >>= 1;
    return r;
}

struct Launch { int maxc; };

}  // namespace

int gemv_max_cols(const b200_ctx *ctx, int type, size_t rb, int64_t K) {
    GemvParams p; size_t smem;
    for (int nc = 4; nc >= 1; nc--) if (plan(ctx, type, (uint32_t)rb, K, nc, 1, 1, p, smem)) return nc;
    return 0;
}

// `gx` describes the activation source: either pre-quantised scratch (act) or f32 (+fusion operands)
struct GemvAct {
    int mode;
    const uint8_t *act;      // ACT_PREQ
    const float *x;          // f32 source, column stride x_stride bytes
    size_t x_stride;
    const float *x2;         // norm weight / up
    float eps;
};

static int gemv_run(b200_ctx *ctx, int type, const uint8_t *W, size_t row_bytes, int64_t N, int64_t K, const GemvAct &ga, int ncols,
                    float *dst, size_t dst_col_stride, const float *residual, bool w_const, int32_t *dbgP, int32_t *dbgM) {
    if (N <= 0 || ncols <= 0) return B200_OK;
    static bool env_read = false;
    if (!env_read) {
        if (const char *e = getenv("GGML_B200_GEMV_WARPS")) g_gemv_warps = atoi(e);
        if (const char *e = getenv("GGML_B200_GEMV_STAGE_KB")) g_gemv_stage_kb = atoi(e);
        if (const char *e = getenv("GGML_B200_GEMV_SMEM_KB")) g_gemv_smem_kb = atoi(e);
        if (const char *e = getenv("GGML_B200_GEMV_CTAS")) g_gemv_ctas = atoi(e);
        env_read = true;
    }
    const int q8k = b200_act_mode_q8k(type);
    const ActLayout L = ActLayout::make(q8k, K);
    if ((type == B200_TYPE_Q4_K || type == B200_TYPE_Q5_K) && (((uintptr_t)W & 15) || (row_bytes & 15))) {
        b200_set_error("gemv: K-quant rows must be 16-byte aligned");
        return B200_ERR_UNSUPPORTED;
    }
    const int maxc = gemv_max_cols(ctx, type, row_bytes, K);
    if (maxc == 0) { b200_set_error("gemv: K=%lld does not fit shared memory", (long long)K); return B200_ERR_UNSUPPORTED; }
    const int nwarps = g_gemv_warps > 0 ? (g_gemv_warps > 8 ? 8 : g_gemv_warps) : 8;
    // one CTA per SM with <= half the shared memory, so that under PDL the NEXT matmul's CTAs co-reside and stream
    // their weights while this one computes; very tall matrices (output projection) take both slots themselves
    const int ctas_per_sm = g_gemv_ctas > 0 ? g_gemv_ctas : (N >= 32768 ? 2 : 1);
    for (int c0 = 0; c0 < ncols; c0 += maxc) {
        const int nc = ncols - c0 < maxc ? ncols - c0 : maxc;
        const int rpw = choose_rpw((uint32_t)row_bytes, nc, N / ((int64_t)ctx->sm_count * ctas_per_sm));
        GemvParams p = {};
        size_t smem = 0;
        int cps = ctas_per_sm;
        if (!plan(ctx, type, (uint32_t)row_bytes, K, nc, rpw, cps, p, smem)) { b200_set_error("gemv: plan failed"); return B200_ERR_FAILED; }
        if (smem > 114 * 1024) cps = 1;
        if (g_gemv_ctas == 0 && cps == 1 && N >= 32768 && smem <= 114 * 1024) cps = 2;
        p.W = W; p.rb = (uint32_t)row_bytes; p.N = (int)N; p.K = (int)K;
        p.act_mode = ga.mode;
        p.act = ga.act ? ga.act + (size_t)c0 * L.col_bytes : nullptr;
        p.L = L;
        p.x = ga.x ? (const float *)((const char *)ga.x + (size_t)c0 * ga.x_stride) : nullptr;
        p.x_stride = ga.x_stride;
        p.x2 = ga.mode == ACT_F32_SWIGLU && ga.x2 ? (const float *)((const char *)ga.x2 + (size_t)c0 * ga.x_stride) : ga.x2;
        p.eps = ga.eps;
        p.dst = dst + (size_t)c0 * dst_col_stride; p.dst_stride = dst_col_stride;
        p.residual = residual ? residual + (size_t)c0 * dst_col_stride : nullptr;
        p.w_const = w_const ? 1 : 0;
        p.use_pdl = ctx->opt_pdl;
        p.dbgP = dbgP; p.dbgM = dbgM;
        p.prof = (unsigned long long *)ctx->prof_buf;
        const int64_t chunks = (N + rpw - 1) / rpw;
        const int64_t maxg = (int64_t)ctx->sm_count * cps;
        const int grid = (int)(chunks < maxg ? chunks : maxg);
        int rc = dbgP ? launch_type<true>(ctx, type, p, nc, rpw, grid, nwarps, smem) : launch_type<false>(ctx, type, p, nc, rpw, grid, nwarps, smem);
        if (rc) return rc;
    }
    return B200_OK;
}

int launch_gemv(b200_ctx *ctx, int type, const uint8_t *W, size_t row_bytes, int64_t N, int64_t K, const uint8_t *act, int ncols,
                float *dst, size_t dst_col_stride, bool w_const) {
    GemvAct ga = {ACT_PREQ, act, nullptr, 0, nullptr, 0.0f};
    return gemv_run(ctx, type, W, row_bytes, N, K, ga, ncols, dst, dst_col_stride, nullptr, w_const, nullptr, nullptr);
}

// fused: dst = W . quant(f(x)) (+ residual); fuse_mode 0: f = id, 1: f = rms_norm(x)*x2, 2: f = silu(x)*x2
int launch_gemv_f32(b200_ctx *ctx, int type, const uint8_t *W, size_t row_bytes, int64_t N, int64_t K, const float *x, size_t x_stride_bytes,
                    int ncols, float *dst, size_t dst_col_stride, bool w_const, int fuse_mode, const float *x2, float eps,
                    const float *residual) {
    GemvAct ga = {fuse_mode == 1 ? ACT_F32_NORM : fuse_mode == 2 ? ACT_F32_SWIGLU : ACT_F32, nullptr, x, x_stride_bytes, x2, eps};
    return gemv_run(ctx, type, W, row_bytes, N, K, ga, ncols, dst, dst_col_stride, residual, w_const, nullptr, nullptr);
}

// test hook: exact integer block sums computed by the SAME item decoders through the SAME pipeline
extern "C" int b200_block_sums(b200_ctx *ctx, int32_t type, const void *W, const float *x, int64_t N, int64_t K, int32_t *P, int32_t *M) {
    if (!ctx || !b200_type_is_quant(type)) { b200_set_error("block_sums: type %d", type); return B200_ERR_UNSUPPORTED; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t nblk = K / b200_type_block_elems(type);
    float *dummy = (float *)ctx->get_scratch(SCRATCH_MISC, (size_t)N * 4);
    if (!dummy) return B200_ERR_ALLOC;
    CUDA_TRY(cudaMemsetAsync(P, 0, (size_t)N * nblk * 4, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(M, 0, (size_t)N * nblk * 4, ctx->stream));
    GemvAct ga = {ACT_F32, nullptr, x, (size_t)K * 4, nullptr, 0.0f};
    return gemv_run(ctx, type, (const uint8_t *)W, b200_row_bytes(type, K), N, K, ga, 1, dummy, (size_t)N, nullptr, false, P, M);
}
