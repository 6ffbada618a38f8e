// dstep.cu -- the persistent decode-step kernel: ONE launch per decoded token.
//
// Round 1 ran a batch-1 llama step as 226 launches (6 fused kernels per layer).  The streaming GEMV itself reached 99% of the
// HBM copy peak on large matrices, but a step averaged 0.51: every kernel boundary drains the memory pipeline (HBM idles
// while the next launch starts, reloads its activations and re-quantises them), and 4096 x 4096 matrices are only 1.4 us of
// HBM time (profiles/r1_gemv_diag.md).  This kernel removes the boundaries instead of shaving them:
//   * one CTA per SM (grid = SM count, cooperative launch), 31 consumer warps + 1 producer warp, alive for the whole step;
//   * the step is a PROGRAM of phases in device memory (DsPhase[]: GEMV over 1-3 weight matrices with a fused rms_norm /
//     silu*up prologue and residual epilogue; rope + KV store + split-KV attention; split merge; row gather), built by
//     graph.cu's matcher from the ggml op list -- the same fused nodes the per-launch kernels take;
//   * weights are constants, so the producer warp streams them through ONE shared-memory ring (cp.async.bulk + mbarrier,
//     L2 evict_first) for the whole program without ever waiting for a phase to finish: while the consumers sit in a grid
//     barrier, re-quantise activations or run attention, the ring (31 stages, ~190 KB per SM = 28 MB chip-wide, 4 us of HBM
//     time) keeps filling with the NEXT matmul's rows.  HBM streams across every dependency of the step;
//   * phases are separated by a grid barrier (one atomic per CTA + acquire spin, ~0.5 us) instead of a kernel boundary
//     (~3-4 us of launch, drain, prologue and tail effects);
//   * long rows (K = 14336: 8-12 KB) are cut into K-pieces of one stage each so that all 31 consumer warps stay busy; a
//     row's pieces are combined in a fixed order at the end of the phase (bit-reproducible run to run);
//   * attention for the single query token runs inside the kernel: CTA (kv head, split) ropes q (and k, v for the split that
//     owns the new cell: converted to the cache type and stored), 28 warps walk the cells with an online softmax, the
//     splits are merged by one CTA per head.
// Arithmetic is that of gemv_bs1.cu / glue.cu / fattn.cu's fast path: the CPU oracle's integers bit for bit (same block
// decoders, gemv_bs1_items.cuh), f32 combination in this kernel's own order.  The cpu-exact mode does not use it.
// Replaces, for a decode step: mul_mat_vec_q + quantize_q8_1 (mmvq.cu:130-204, quantize.cu:4-38), rms_norm_f32, rope_norm,
// cpy_f32_f16 / cpy_blck_f32_q8_0, flash_attn_vec_ext_f32 + flash_attn_combine_results, k_bin_bcast add, silu
// (norm.cu, rope.cu, cpy.cu, fattn-vec-f32.cuh, fattn-common.cuh:609-650, binbcast.cu, unary.cu) and the CUDA-graph node
// chain of ggml_backend_cuda_graph_compute (ggml-cuda.cu:2432-2788).
#include "dstep.h"
#include "quant_warp.cuh"
#include "gemv_bs1_items.cuh"
#include <string.h>
#include <algorithm>

namespace {

using namespace bs1;

constexpr int DS_NCW = 31;                 // consumer warps
constexpr int DS_THREADS = (DS_NCW + 1) * 32;
constexpr int DS_MAX_STAGES = 31;          // one producer lane per stage, stage s is consumed by warp s
constexpr int DS_PART_ROWS = 128;          // rows per CTA of a K-split segment
constexpr int DS_MAX_PIECES = 4;
constexpr int TB_Q4_K = 1, TB_Q5_K = 2, TB_Q6_K = 4;
enum { KVT_F16 = 0, KVT_Q8_0 = 1, KVT_Q4_0 = 2 };

struct DsSeg {
    const uint8_t *W;
    float *        dst;
    const float *  residual;
    uint32_t       rb;                     // bytes per row
    int            type, N;
    int            R, lgR;                 // rows per chunk (power of two; 1 when the row is K-split)
    int            S, nbp;                 // K-pieces per row (1..4) and 256-blocks per piece
    int            q, rem;                 // N = q * grid + rem: CTA c owns rows [c*q + min(c, rem), ...)
};
struct DsGemv {
    int            nseg, K, act_mode;
    float          eps;
    const float *  x, *x2;
    uint32_t       off_aq64, off_aq128, off_ad, off_s32, off_s16;     // activation layouts of THIS phase inside the activation area (0 = not needed: aq*, s*)
    DsSeg          seg[GEMV_MAX_SEG];
};
struct DsAttn {
    const float *  q, *k, *v;              // raw q/k/v of the token (f32, [H*D], [Hkv*D], [Hkv*D]) from the qkv phase
    const int32_t *pos;
    const float *  ff;                     // rope frequency factors (optional)
    const char *   kc, *vc;                // cache views of the FLASH_ATTN_EXT op: cell stride nb1, head stride nb2
    uint64_t       k_nb1, k_nb2, v_nb1, v_nb2;
    const char *   mask;                   // f16 [n_kv] row of the token
    void *         k_dst, *v_dst;          // cache rows of the new token ([Hkv][D] in the cache type) ...
    void *const *  k_dst_ind, *const *v_dst_ind;      // ... or where to read them from (CUDA-graph replay, graph.cu)
    float *        part;                   // [Hkv * nsplit][gq][D + 2] split partials
    float *        out;                    // attention output [H * D]
    int            H, Hkv, gq, n_kv, kvt, nsplit, len, npw;
    float          scale;
    RopeParams     rp;
};
struct DsCopy { const char *src; const int32_t *idx; const float *src2; uint64_t nb1; float *dst; int ne0, nrows, add; };   // get_rows (f32 rows), or dst = src + src2
struct DsPhase {
    int kind, pad;
    union { DsGemv g; DsAttn a; DsCopy c; };
};

struct DsParams {
    const DsPhase *prog;
    int            nphases;
    unsigned int * sync;                   // [0] barrier arrivals (monotonic within a launch), [1] exits, [2] error flag
    int            nstages;
    uint32_t       stage_bytes;
    uint32_t       off_part, off_desc, off_pgeo, off_act, off_ring;
    unsigned long long *prof;              // debug: [grid][nphases][4] %globaltimer stamps: phase start, prologue done, work done (warp 0), barrier passed
};

__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ldcg_f(const float *p) { return __ldcg(p); }

// per-CTA geometry of a GEMV phase, needed by the producer and the consumers alike.  Only ever indexed with compile-time
// constants (unrolled loops + selects): it must live in registers -- with ~210 KB of shared memory per CTA there is next to no L1
// left, and a stack frame would turn every access into an L2 round trip.
struct PhaseGeo { int lo[GEMV_MAX_SEG], hi[GEMV_MAX_SEG], ch0[GEMV_MAX_SEG + 1], poff[GEMV_MAX_SEG], pad[3]; };
static_assert(sizeof(PhaseGeo) == 64, "PhaseGeo");
__device__ __forceinline__ int phase_geo(const DsGemv &g, int c, PhaseGeo *o) {       // o: shared memory; returns the chunk count
    int po = 0, ch = 0;
    o->ch0[0] = 0;
#pragma unroll
    for (int s = 0; s < GEMV_MAX_SEG; s++) {
        int lo = 0, hi = 0, n = 0;
        const int po_s = po;
        if (s < g.nseg) {
            const int q = g.seg[s].q, rem = g.seg[s].rem, S = g.seg[s].S;
            lo = c * q + min(c, rem);
            hi = (c + 1) * q + min(c + 1, rem);
            const int rows = hi - lo;
            n = ((rows + g.seg[s].R - 1) >> g.seg[s].lgR) * S;
            if (S > 1) po += rows;
        }
        ch += n;
        o->lo[s] = lo; o->hi[s] = hi; o->poff[s] = po_s; o->ch0[s + 1] = ch;
    }
    return ch;
}

// grid-wide barrier between phases: consumers only (the producer never stops streaming).  Pattern of cooperative groups'
// grid.sync: CTA barrier, one thread fences + arrives + spins with acquire loads, fence, CTA barrier.
__device__ __forceinline__ void grid_sync(const DsParams &P, int &nbar) {
    named_bar_sync(1, DS_NCW * 32);
    nbar++;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&P.sync[0], 1u);
        const unsigned target = (unsigned)nbar * gridDim.x;
        const long long t0 = clock64();
        while (ld_acquire_u32(&P.sync[0]) < target) {
            if (clock64() - t0 > (1ll << 33)) { atomicExch(&P.sync[2], 1u); break; }        // ~4 s: a CTA is missing (never with a cooperative launch)
        }
        __threadfence();
    }
    named_bar_sync(1, DS_NCW * 32);
}

// ---------------------------------------------------------------------------------------------- GEMV phase: activation prologue
struct ActOff { uint32_t aq64, aq128, ad, s32, s16; bool n64, n128; };      // byte offsets from the CTA's shared memory base
__device__ __forceinline__ void quant_block_to_smem(const ActOff &A, uint8_t *smem, int b, int lane, const float (&v)[8]) {
    const int e0 = b * 256 + lane * 8;
    uint2 qp; float d; int pair;
    warp_quant_q8k(v, lane, qp, d, pair);
    const int quad = pair + __shfl_xor_sync(0xffffffffu, pair, 2);
    if (A.n128 && (lane & 1) == 0) *(int16_t *)(smem + A.s16 + (uint32_t)(b * 16 + (lane >> 1)) * 2) = (int16_t)pair;
    if (A.n64 && (lane & 3) == 0) *(int16_t *)(smem + A.s32 + (uint32_t)(b * 8 + (lane >> 2)) * 2) = (int16_t)quad;
    if (lane == 0) *(float *)(smem + A.ad + (uint32_t)b * 4) = d;
    if (A.n64)  *(uint2 *)(smem + A.aq64 + (uint32_t)((e0 >> 8) * 272 + (e0 & 255))) = qp;
    if (A.n128) *(uint2 *)(smem + A.aq128 + (uint32_t)((e0 >> 7) * 144 + (e0 & 127))) = qp;
}
__device__ __forceinline__ void apply_mode(int mode, float (&v)[8], const float4 &w0, const float4 &w1, float norm_scale) {
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    if (mode == ACT_F32_NORM) {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __fmul_rn(__fmul_rn(v[j], norm_scale), w[j]);
    } else if (mode == ACT_F32_SWIGLU) {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __fmul_rn(ggml_silu_lane(v[j]), w[j]);
    }
}
__device__ __forceinline__ float norm_scale_of(double *s_red, double s, int K, float eps, int warp, int lane) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_red[warp] = s;
    named_bar_sync(2, DS_NCW * 32);
    double t = lane < DS_NCW ? s_red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    named_bar_sync(2, DS_NCW * 32);                 // s_red is reused by the next phase
    return __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn((float)(t / (double)K), eps)));
}
__device__ __forceinline__ float4 ld4cg(const float *p) { return __ldcg((const float4 *)p); }
// f32 activations (written by other CTAs in an earlier phase: read through L2) -> q8_K in shared memory, bit-exact vs
// quantize_row_q8_K_ref; optional rms_norm(x)*w or silu(g)*u first
__device__ __forceinline__ void gemv_prologue(const ActOff &P, const DsGemv &g, uint8_t *smem, double *s_red, int warp, int lane) {
    const int nchunk = g.K >> 8, mode = g.act_mode;
    if (nchunk <= 2 * DS_NCW) {
        float4 xa[2][2], xb[2][2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int b = warp + u * DS_NCW;
            xa[u][0] = xa[u][1] = xb[u][0] = xb[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b < nchunk) {
                const int e0 = b * 256 + lane * 8;
                xa[u][0] = ld4cg(g.x + e0); xa[u][1] = ld4cg(g.x + e0 + 4);
                if (mode == ACT_F32_NORM) { xb[u][0] = *(const float4 *)(g.x2 + e0); xb[u][1] = *(const float4 *)(g.x2 + e0 + 4); }
                else if (mode == ACT_F32_SWIGLU) { xb[u][0] = ld4cg(g.x2 + e0); xb[u][1] = ld4cg(g.x2 + e0 + 4); }
            }
        }
        float ns = 1.0f;
        if (mode == ACT_F32_NORM) {
            double s = 0.0;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const float v[8] = {xa[u][0].x, xa[u][0].y, xa[u][0].z, xa[u][0].w, xa[u][1].x, xa[u][1].y, xa[u][1].z, xa[u][1].w};
#pragma unroll
                for (int j = 0; j < 8; j++) s += (double)__fmul_rn(v[j], v[j]);
            }
            ns = norm_scale_of(s_red, s, g.K, g.eps, warp, lane);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int b = warp + u * DS_NCW;
            if (b < nchunk) {
                float v[8] = {xa[u][0].x, xa[u][0].y, xa[u][0].z, xa[u][0].w, xa[u][1].x, xa[u][1].y, xa[u][1].z, xa[u][1].w};
                apply_mode(mode, v, xb[u][0], xb[u][1], ns);
                quant_block_to_smem(P, smem, b, lane, v);
            }
        }
        return;
    }
    float ns = 1.0f;
    if (mode == ACT_F32_NORM) {
        double s = 0.0;
        for (int i = threadIdx.x; i < g.K; i += DS_NCW * 32) { const float v = ldcg_f(g.x + i); s += (double)__fmul_rn(v, v); }
        ns = norm_scale_of(s_red, s, g.K, g.eps, warp, lane);
    }
#pragma unroll 1
    for (int b = warp; b < nchunk; b += DS_NCW) {
        const int e0 = b * 256 + lane * 8;
        const float4 a0 = ld4cg(g.x + e0), a1 = ld4cg(g.x + e0 + 4);
        float4 w0 = a0, w1 = a1;
        if (mode == ACT_F32_NORM) { w0 = *(const float4 *)(g.x2 + e0); w1 = *(const float4 *)(g.x2 + e0 + 4); }
        else if (mode == ACT_F32_SWIGLU) { w0 = ld4cg(g.x2 + e0); w1 = ld4cg(g.x2 + e0 + 4); }
        float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        apply_mode(mode, v, w0, w1, ns);
        quant_block_to_smem(P, smem, b, lane, v);
    }
}

// ---------------------------------------------------------------------------------------------- attention phase helpers
__device__ __forceinline__ uint32_t ld_u16x2_cg(const uint8_t *p) {         // 4 bytes at a 2-byte aligned address, through L2
    return (uint32_t)__ldcg((const unsigned short *)p) | ((uint32_t)__ldcg((const unsigned short *)(p + 2)) << 16);
}
// rope of one pair, arithmetic identical to glue.cu's b200_rope_store_kernel (fast mode)
__device__ __forceinline__ void rope_pair(const RopeParams &rp, const float *ff, int p, int ip, const float *src, float *dst) {
    const int i0 = 2 * ip;
    if (i0 < rp.n_dims) {
        float theta = (float)p;
        for (int k = 0; k < ip; k++) theta = __fmul_rn(theta, rp.theta_scale);
        const float f = ff ? ff[ip] : 1.0f;
        const float te = __fdiv_rn(theta, f);
        float ti = __fmul_rn(rp.freq_scale, te), th = ti, ms = rp.attn_factor;
        if (rp.ext_factor != 0.0f) {
            const float yv = __fdiv_rn((float)(i0 / 2) - rp.corr0, fmaxf(0.001f, rp.corr1 - rp.corr0));
            const float ramp = __fmul_rn(1.0f - fminf(1.0f, fmaxf(0.0f, yv)), rp.ext_factor);
            th = __fadd_rn(__fmul_rn(ti, 1.0f - ramp), __fmul_rn(te, ramp));
            ms = __fmul_rn(ms, 1.0f + 0.1f * logf(__fdiv_rn(1.0f, rp.freq_scale)));
        }
        const float c = __fmul_rn(cosf(th), ms), s = __fmul_rn(sinf(th), ms);
        const bool neox = (rp.mode & 2) != 0;
        const int ia = neox ? ip : i0, ib = neox ? ip + rp.n_dims / 2 : i0 + 1;
        const float x0 = ldcg_f(src + ia), x1 = ldcg_f(src + ib);
        dst[ia] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
        dst[ib] = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
    } else {
        dst[i0] = ldcg_f(src + i0);
        dst[i0 + 1] = ldcg_f(src + i0 + 1);
    }
}
// KV-store conversion of one 128-float row in shared memory to the cache type (glue.cu semantics: f16 RNE,
// quantize_row_q8_0 AVX2 path, quantize_row_q4_0_ref); called by 128 threads with t = 0..127
__device__ __forceinline__ void store_kv_row(int kvt, const float *row, void *dst, int t) {
    constexpr int D = 128;
    if (kvt == KVT_F16) { ((__half *)dst)[t] = __float2half_rn(row[t]); return; }
    if (t >= D / 32) return;
    const float *v = row + t * 32;
    if (kvt == KVT_Q8_0) {
        uint8_t *o = (uint8_t *)dst + t * 34;
        float amax = 0.0f;
        for (int j = 0; j < 32; j++) amax = fmaxf(amax, fabsf(v[j]));
        const float dd = __fdiv_rn(amax, 127.0f);
        const float id = amax != 0.0f ? __fdiv_rn(127.0f, amax) : 0.0f;
        *(__half *)o = __float2half_rn(dd);
        for (int j = 0; j < 32; j++) o[2 + j] = (uint8_t)(int8_t)__float2int_rn(__fmul_rn(v[j], id));
    } else {
        uint8_t *o = (uint8_t *)dst + t * 18;
        float amax = 0.0f, mx = 0.0f;
        for (int j = 0; j < 32; j++) { const float a = fabsf(v[j]); if (a > amax) { amax = a; mx = v[j]; } }
        const float dd = __fdiv_rn(mx, -8.0f);
        const float id = dd != 0.0f ? __fdiv_rn(1.0f, dd) : 0.0f;
        *(__half *)o = __float2half_rn(dd);
        for (int j = 0; j < 16; j++) {
            const float x0 = __fmul_rn(v[j], id), x1 = __fmul_rn(v[16 + j], id);
            int a = (int)(int8_t)(int)__fadd_rn(x0, 8.5f), c = (int)(int8_t)(int)__fadd_rn(x1, 8.5f);
            a = a > 15 ? 15 : a; c = c > 15 ? 15 : c;
            o[2 + j] = (uint8_t)((a & 0xff) | (c << 4));
        }
    }
}

// one (kv head, split) unit: rope, KV store (owner split), online-softmax walk over the cells, merge of the CTA's warps
__device__ __forceinline__ void attn_unit(const DsAttn &a, uint8_t *scr, int warp, int lane) {
    constexpr int D = 128;
    const int unit = blockIdx.x;
    const bool has_unit = unit < a.Hkv * a.nsplit;
    const int hk = has_unit ? unit / a.nsplit : 0, sp = has_unit ? unit % a.nsplit : 0;
    const int c0 = sp * a.len, c1 = min(a.n_kv, c0 + a.len);
    float *sq = (float *)scr;                         // [gq][D] roped q
    float *skv = sq + a.gq * D;                       // [2][D]: roped k, v of the new token (owner split only)
    float *mbuf = skv + 2 * D;                        // [gq * npw][MB] per-warp partials (MB = D + 4: rows stay 16-byte aligned)
    constexpr int MB = D + 4;
    const int tid = threadIdx.x;
    if (!has_unit) return;
    const int p = __ldcg(a.pos);
    char *kd = (char *)(a.k_dst_ind ? *a.k_dst_ind : a.k_dst), *vd = (char *)(a.v_dst_ind ? *a.v_dst_ind : a.v_dst);
    const int cell = (int)((kd - a.kc) / (long long)a.k_nb1);
    const bool owner = cell >= c0 && cell < c1;
    // ---- rope: gq query heads (+ the k head), one pair per thread ----
    for (int t = tid; t < (a.gq + 1) * (D / 2); t += DS_NCW * 32) {
        const int h = t / (D / 2), ip = t % (D / 2);
        if (h < a.gq) rope_pair(a.rp, a.ff, p, ip, a.q + (size_t)(hk * a.gq + h) * D, sq + h * D);
        else if (owner) rope_pair(a.rp, a.ff, p, ip, a.k + (size_t)hk * D, skv);
    }
    if (owner) for (int t = tid; t < D; t += DS_NCW * 32) skv[D + t] = ldcg_f(a.v + (size_t)hk * D + t);
    named_bar_sync(3, DS_NCW * 32);
    if (owner) {
        const size_t rowb = a.kvt == KVT_F16 ? (size_t)D * 2 : a.kvt == KVT_Q8_0 ? (size_t)(D / 32) * 34 : (size_t)(D / 32) * 18;
        if (tid < D) store_kv_row(a.kvt, skv, kd + (size_t)hk * rowb, tid);
        else if (tid < 2 * D) store_kv_row(a.kvt, skv + D, vd + (size_t)hk * rowb, tid - D);
        __threadfence_block();
    }
    named_bar_sync(3, DS_NCW * 32);                   // the new cell is in the cache (this CTA reads it back through L2)
    // ---- walk the cells: warp = (head of the group, interleave) ----
    const int hg = warp % a.gq, sub = warp / a.gq;
    const bool walker = sub < a.npw;
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f, mrow = -INFINITY, lrow = 0.0f;
    if (walker) {
        const float4 qf = *(const float4 *)(sq + hg * D + lane * 4);
        float q0 = qf.x, q1 = qf.y, q2 = qf.z, q3 = qf.w, dq = 0.0f;
        int qi = 0;
        if (a.kvt == KVT_F16) {        // the CPU rounds Q to f16 for an f16 K (vec_dot_type)
            q0 = __half2float(__float2half_rn(q0)); q1 = __half2float(__float2half_rn(q1));
            q2 = __half2float(__float2half_rn(q2)); q3 = __half2float(__float2half_rn(q3));
        } else {                       // ... and quantises it to q8_0 for a quantised K: block = 8 lanes x 4 elements
            float amax = fmaxf(fmaxf(fabsf(q0), fabsf(q1)), fmaxf(fabsf(q2), fabsf(q3)));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
            const float d = __fdiv_rn(amax, 127.0f), id = amax != 0.0f ? __fdiv_rn(127.0f, amax) : 0.0f;
            dq = __half2float(__float2half_rn(d));
            const int a0 = __float2int_rn(__fmul_rn(q0, id)), a1 = __float2int_rn(__fmul_rn(q1, id));
            const int a2 = __float2int_rn(__fmul_rn(q2, id)), a3 = __float2int_rn(__fmul_rn(q3, id));
            qi = (a0 & 0xff) | ((a1 & 0xff) << 8) | ((a2 & 0xff) << 16) | ((a3 & 0xff) << 24);
        }
        const char *kb = a.kc + (size_t)hk * a.k_nb2, *vb = a.vc + (size_t)hk * a.v_nb2;
        const int blk = lane >> 3, e0 = (lane & 7) * 4;        // quantised rows: 32-element block and offset inside it
#pragma unroll 2
        for (int c = c0 + sub; c < c1; c += a.npw) {
            const __half mh = __ldcg((const __half *)(a.mask + (size_t)c * 2));
            const float mv = __half2float(mh);
            if (__hisinf(mh) && mv < 0.0f) continue;
            const uint8_t *kr = (const uint8_t *)(kb + (size_t)c * a.k_nb1), *vr = (const uint8_t *)(vb + (size_t)c * a.v_nb1);
            float s, v0, v1, v2, v3;
            if (a.kvt == KVT_F16) {
                const uint2 kk = __ldcg((const uint2 *)(kr + lane * 8)), vv = __ldcg((const uint2 *)(vr + lane * 8));
                const float2 k01 = __half22float2(*(const __half2 *)&kk.x), k23 = __half22float2(*(const __half2 *)&kk.y);
                const float2 v01 = __half22float2(*(const __half2 *)&vv.x), v23 = __half22float2(*(const __half2 *)&vv.y);
                s = fmaf(q3, k23.y, fmaf(q2, k23.x, fmaf(q1, k01.y, __fmul_rn(q0, k01.x))));
                s = warp_reduce_sum(s);
                v0 = v01.x; v1 = v01.y; v2 = v23.x; v3 = v23.y;
            } else {
                const int bb = a.kvt == KVT_Q8_0 ? 34 : 18;
                const uint8_t *kblk = kr + blk * bb, *vblk = vr + blk * bb;
                const float dk = __half2float(__ushort_as_half(__ldcg((const unsigned short *)kblk)));
                const float dv = __half2float(__ushort_as_half(__ldcg((const unsigned short *)vblk)));
                uint32_t kw, vw;
                if (a.kvt == KVT_Q8_0) { kw = ld_u16x2_cg(kblk + 2 + e0); vw = ld_u16x2_cg(vblk + 2 + e0); }
                else {
                    const int o = 2 + (e0 & 15);
                    const uint32_t kraw = ld_u16x2_cg(kblk + o), vraw = ld_u16x2_cg(vblk + o);
                    kw = __vsub4(e0 < 16 ? (kraw & 0x0f0f0f0fu) : ((kraw >> 4) & 0x0f0f0f0fu), 0x08080808u);
                    vw = __vsub4(e0 < 16 ? (vraw & 0x0f0f0f0fu) : ((vraw >> 4) & 0x0f0f0f0fu), 0x08080808u);
                }
                int isum = __dp4a((int)kw, qi, 0);
                isum += __shfl_xor_sync(0xffffffffu, isum, 1);
                isum += __shfl_xor_sync(0xffffffffu, isum, 2);
                isum += __shfl_xor_sync(0xffffffffu, isum, 4);
                s = __fmul_rn((float)isum, __fmul_rn(dk, dq));                 // every lane of a block holds the block's term
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                v0 = __fmul_rn((float)(int8_t)(vw & 0xff), dv); v1 = __fmul_rn((float)(int8_t)((vw >> 8) & 0xff), dv);
                v2 = __fmul_rn((float)(int8_t)((vw >> 16) & 0xff), dv); v3 = __fmul_rn((float)(int8_t)(vw >> 24), dv);
            }
            s = fmaf(s, a.scale, mv);
            const float mnew = fmaxf(mrow, s);
            const float corr = mrow == -INFINITY ? 0.0f : expf(mrow - mnew), pw = expf(s - mnew);
            mrow = mnew;
            lrow = fmaf(lrow, corr, pw);
            o0 = fmaf(o0, corr, pw * v0); o1 = fmaf(o1, corr, pw * v1); o2 = fmaf(o2, corr, pw * v2); o3 = fmaf(o3, corr, pw * v3);
        }
        float *mb = mbuf + (size_t)(hg * a.npw + sub) * MB;
        *(float4 *)(mb + lane * 4) = make_float4(o0, o1, o2, o3);
        if (lane == 0) { mb[D] = mrow; mb[D + 1] = lrow; }
    }
    named_bar_sync(3, DS_NCW * 32);
    // ---- merge the interleaves of every head (fixed order) and write the split partial ----
    if (warp < a.gq) {
        const float *mb = mbuf + (size_t)warp * a.npw * MB;
        float M = -INFINITY;
        for (int i = 0; i < a.npw; i++) M = fmaxf(M, mb[i * MB + D]);
        float L = 0.0f, r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
        if (M != -INFINITY) {
            for (int i = 0; i < a.npw; i++) {
                const float mi = mb[i * MB + D];
                const float f = mi == -INFINITY ? 0.0f : expf(mi - M);
                const float4 xx = *(const float4 *)(mb + i * MB + lane * 4);
                L = fmaf(mb[i * MB + D + 1], f, L);
                r0 = fmaf(xx.x, f, r0); r1 = fmaf(xx.y, f, r1); r2 = fmaf(xx.z, f, r2); r3 = fmaf(xx.w, f, r3);
            }
        }
        float *pp = a.part + ((size_t)unit * a.gq + warp) * (D + 2);
        *(float2 *)(pp + lane * 4) = make_float2(r0, r1);
        *(float2 *)(pp + lane * 4 + 2) = make_float2(r2, r3);
        if (lane == 0) { pp[D] = M; pp[D + 1] = L; }
    }
}

// merge of the KV splits of one head: CTA h < H, warps 0..3 take 32 dims each
__device__ __forceinline__ void attn_combine(const DsAttn &a, int warp, int lane) {
    constexpr int D = 128;
    const int h = blockIdx.x;
    if (h >= a.H || warp >= 4) return;
    const int hk = h / a.gq, hg = h % a.gq;
    const float *base = a.part + ((size_t)(hk * a.nsplit) * a.gq + hg) * (D + 2);
    const size_t sstride = (size_t)a.gq * (D + 2);
    float M = -INFINITY;
    for (int s = 0; s < a.nsplit; s++) M = fmaxf(M, ldcg_f(base + s * sstride + D));
    float L = 0.0f, acc = 0.0f;
    const int d = warp * 32 + lane;
    for (int s = 0; s < a.nsplit; s++) {
        const float ms = ldcg_f(base + s * sstride + D);
        const float f = ms == -INFINITY ? 0.0f : expf(ms - M);
        L = fmaf(ldcg_f(base + s * sstride + D + 1), f, L);
        acc = fmaf(ldcg_f(base + s * sstride + d), f, acc);
    }
    a.out[(size_t)h * D + d] = acc / L;
}

// ---------------------------------------------------------------------------------------------- the kernel
constexpr int DS_DESC_BYTES = 320;
static_assert(sizeof(DsPhase) <= DS_DESC_BYTES && sizeof(DsPhase) % 16 == 0, "phase descriptor size");

template <int TYPES>
__global__ void __launch_bounds__(DS_THREADS, 1) b200_decode_step_kernel(const DsParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full  = (uint64_t *)smem;                      // [32]
    uint64_t *empty = full + 32;                             // [32]
    double *  s_red = (double *)(empty + 32);                // [32]
    float *   part  = (float *)(smem + P.off_part);          // [DS_MAX_PIECES][DS_PART_ROWS]
    uint8_t * ring  = smem + P.off_ring;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x, ns = P.nstages;

    if (warp == DS_NCW) {
        // ================================================================== producer: lane l owns ring stage l, for the whole step
        if (lane < ns) { mbar_init(&full[lane], 1); mbar_init(&empty[lane], 1); }
        mbar_fence_init();
        __syncwarp();
        asm volatile("bar.arrive 4, %0;" ::"r"(DS_THREADS) : "memory");
        const uint64_t pol = l2_policy_evict_first();
        bool active = lane < ns;
        int ph = -1, gi = lane, cb = 0, ce = 0, use = 0;         // gi: my next chunk (global index); [cb, ce): chunk range of phase ph
        PhaseGeo *geo = (PhaseGeo *)(smem + P.off_pgeo) + lane;          // each lane may be in a different phase: its own copy
        while (__any_sync(0xffffffffu, active)) {
            if (active) {
                while (gi >= ce) {
                    ph++;
                    if (ph >= P.nphases) { active = false; break; }
                    if (P.prog[ph].kind == DS_GEMV) { cb = ce; ce = cb + phase_geo(P.prog[ph].g, c, geo); }
                }
                if (active && (use == 0 || mbar_test_wait(&empty[lane], (use - 1) & 1))) {
                    const DsGemv &g = P.prog[ph].g;
                    const int li = gi - cb;
                    int sidx = 0;
#pragma unroll
                    for (int t = 1; t < GEMV_MAX_SEG; t++) if (t < g.nseg && li >= geo->ch0[t]) sidx = t;
                    const DsSeg &sg = g.seg[sidx];
                    const int S = sg.S, R = sg.R, lgR = sg.lgR, nbp = sg.nbp, ty = sg.type;
                    const uint32_t rb = sg.rb;
                    const int t = li - geo->ch0[sidx], rg = S > 1 ? t / S : t, pc = S > 1 ? t - rg * S : 0;
                    const int row = geo->lo[sidx] + (rg << lgR);
                    const int nr = min(R, geo->hi[sidx] - row);
                    const int nb = g.K >> 8;
                    const uint32_t bbytes = ty == B200_TYPE_Q4_K ? 144u : ty == B200_TYPE_Q5_K ? 176u : 210u;
                    const uint8_t *src = sg.W + (size_t)row * rb + (size_t)(pc * nbp) * bbytes;
                    const uint32_t len = S > 1 ? (uint32_t)min(nbp, nb - pc * nbp) * bbytes : (uint32_t)nr * rb;
                    const uint32_t extra = (uint32_t)((uintptr_t)src & 15);
                    const uint32_t bytes = (extra + len + 15u) & ~15u;
                    mbar_arrive_expect_tx(&full[lane], bytes);
                    bulk_g2s_hint(ring + (size_t)lane * P.stage_bytes, src - extra, bytes, &full[lane], pol);
                    gi += ns; use++;
                }
            }
        }
        return;
    }

    // ====================================================================== consumers
    asm volatile("bar.sync 4, %0;" ::"r"(DS_THREADS) : "memory");        // mbarriers are initialised
    int nbar = 0, use = 0;
    int cbm = 0;                                                          // (global chunk index where the current phase starts) mod ns
    const DsPhase &phd = *(const DsPhase *)(smem + P.off_desc);           // the current phase's descriptor, staged in shared memory
    PhaseGeo *sgeo = (PhaseGeo *)(smem + P.off_desc + DS_DESC_BYTES);     // ... and this CTA's row / chunk geometry of it
#define DSPROF(slot) do { if (P.prof && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); P.prof[((size_t)blockIdx.x * P.nphases + ph) * 4 + (slot)] = t_; } } while (0)
    for (int ph = 0; ph < P.nphases; ph++) {
        DSPROF(0);
        // every consumer reads the descriptor dozens of times per chunk: one 320-byte copy per phase instead of L2 round trips
        if (threadIdx.x < DS_DESC_BYTES / 16) ((uint4 *)(smem + P.off_desc))[threadIdx.x] = __ldg((const uint4 *)&P.prog[ph] + threadIdx.x);
        if (threadIdx.x == 32 && P.prog[ph].kind == DS_GEMV) phase_geo(P.prog[ph].g, c, sgeo);
        named_bar_sync(1, DS_NCW * 32);
        if (phd.kind == DS_GEMV) {
            const DsGemv &g = phd.g;
            ActOff A;
            A.n64 = g.off_aq64 != 0xffffffffu; A.n128 = g.off_aq128 != 0xffffffffu;
            A.aq64 = P.off_act + g.off_aq64; A.aq128 = P.off_act + g.off_aq128; A.ad = P.off_act + g.off_ad; A.s32 = P.off_act + g.off_s32; A.s16 = P.off_act + g.off_s16;
            gemv_prologue(A, g, smem, s_red, warp, lane);
            named_bar_sync(1, DS_NCW * 32);
            DSPROF(1);
            const int nch = sgeo->ch0[GEMV_MAX_SEG], nseg = g.nseg;
            const uint8_t *aq64 = smem + A.aq64, *aq128 = smem + A.aq128;
            const U4 *sums4 = (const U4 *)(smem + A.s32);
            const U4 *s16 = (const U4 *)(smem + A.s16);
            const float *ad = (const float *)(smem + A.ad);
            const int nb = g.K >> 8;
            // my chunks: local index li with (phase start + li) % ns == warp
            int li = warp - cbm; if (li < 0) li += ns;
            if (warp >= ns) li = nch;
#pragma unroll 1
            for (; li < nch; li += ns) {
                int sidx = 0;
#pragma unroll
                for (int t = 1; t < GEMV_MAX_SEG; t++) if (t < nseg && li >= sgeo->ch0[t]) sidx = t;
                const DsSeg &sg = g.seg[sidx];
                const int S = sg.S, R = sg.R, ty = sg.type;
                const uint32_t rb = sg.rb;
                const int t = li - sgeo->ch0[sidx], rg = S > 1 ? t / S : t, pc = S > 1 ? t - rg * S : 0;
                const int row0 = sgeo->lo[sidx] + (rg << sg.lgR);
                const int nr = min(R, sgeo->hi[sidx] - row0);
                const int b0 = pc * sg.nbp, nbk = S > 1 ? min(sg.nbp, nb - b0) : nb;
                const uint32_t bbytes = ty == B200_TYPE_Q4_K ? 144u : ty == B200_TYPE_Q5_K ? 176u : 210u;
                const uint32_t extra = (uint32_t)((uintptr_t)(sg.W + (size_t)row0 * rb + (size_t)b0 * bbytes) & 15);
                const uint8_t *rowp = ring + (size_t)warp * P.stage_bytes + extra;
                // K-split rows land in part[piece][row - lo]; whole rows go straight to dst (+ residual); all read back from the staged descriptor
                const int prt_off = pc * DS_PART_ROWS + sgeo->poff[sidx] - sgeo->lo[sidx];
                mbar_wait(&full[warp], use & 1);
                if ((TYPES & (TB_Q4_K | TB_Q5_K)) && ((TYPES & TB_Q6_K) == 0 || ty != B200_TYPE_Q6_K)) {
                    const bool q5 = (TYPES & TB_Q5_K) && (TYPES == TB_Q5_K || ty == B200_TYPE_Q5_K);
                    const int lpr = R == 1 ? 32 : 16, rpp = 32 / lpr;
                    const int sub = lane / lpr, bl = lane % lpr;
#pragma unroll 1
                    for (int r = 0; r < nr; r += rpp) {
                        const bool mine = r + sub < nr;
                        const uint8_t *rp = rowp + (size_t)(r + (mine ? sub : 0)) * rb;
                        float acc = 0.0f;
#pragma unroll 1
                        for (int blk = bl; blk < nbk; blk += lpr) {
                            const uint8_t *b = rp + blk * bbytes;
                            const int ab = b0 + blk;
                            if ((TYPES & TB_Q5_K) && q5) acc += block_q45k<true>(b, aq64 + ab * 272, sums4[ab], ad[ab]);
                            else if (TYPES & TB_Q4_K) acc += block_q45k<false>(b, aq64 + ab * 272, sums4[ab], ad[ab]);
                        }
                        if (r + rpp >= nr) { __syncwarp(); if (lane == 0) mbar_arrive(&empty[warp]); }
                        if (lpr == 32) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                        if (bl == 0 && mine) {
                            const int o = row0 + r + sub;
                            if (S > 1) part[prt_off + o] = acc;
                            else { const float *residual = sg.residual; if (residual) acc = __fadd_rn(acc, ldcg_f(residual + o)); sg.dst[o] = acc; }
                        }
                    }
                } else if (TYPES & TB_Q6_K) {
                    const int nit = 2 * nbk;
#pragma unroll 1
                    for (int r = 0; r < nr; r++, rowp += rb) {
                        float acc = 0.0f;
#pragma unroll 1
                        for (int it = lane; it < nit; it += 32) acc += item_q6k(rowp, it, aq128 + (size_t)b0 * 288, s16 + 2 * b0, ad + b0);
                        if (r == nr - 1) { __syncwarp(); if (lane == 0) mbar_arrive(&empty[warp]); }
                        acc = warp_reduce_sum(acc);
                        if (lane == 0) {
                            const int o = row0 + r;
                            if (S > 1) part[prt_off + o] = acc;
                            else { const float *residual = sg.residual; if (residual) acc = __fadd_rn(acc, ldcg_f(residual + o)); sg.dst[o] = acc; }
                        }
                    }
                }
                use++;
            }
            // ---- K-split segments: add a row's pieces in piece order, residual, store ----
            bool any_split = false;
#pragma unroll
            for (int s = 0; s < GEMV_MAX_SEG; s++) any_split |= s < nseg && g.seg[s].S > 1;
            if (any_split) {
                named_bar_sync(1, DS_NCW * 32);
#pragma unroll
                for (int s = 0; s < GEMV_MAX_SEG; s++) {
                    if (s < nseg && g.seg[s].S > 1) {
                        const int S = g.seg[s].S, lo = sgeo->lo[s], rows = sgeo->hi[s] - sgeo->lo[s], po = sgeo->poff[s];
                        const float *residual = g.seg[s].residual;
                        float *dst = g.seg[s].dst;
                        for (int r = threadIdx.x; r < rows; r += DS_NCW * 32) {
                            float v = part[po + r];
                            for (int pc = 1; pc < S; pc++) v = __fadd_rn(v, part[pc * DS_PART_ROWS + po + r]);
                            if (residual) v = __fadd_rn(v, ldcg_f(residual + lo + r));
                            dst[lo + r] = v;
                        }
                    }
                }
            }
            cbm = (cbm + nch) % ns;
        } else if (phd.kind == DS_ATTN) {
            attn_unit(phd.a, smem + P.off_act, warp, lane);
        } else if (phd.kind == DS_COMBINE) {
            attn_combine(phd.a, warp, lane);
        } else {        // DS_COPY: dst[r][:] = src[idx[r]][:] (f32 rows), or dst = src + src2
            const DsCopy &cp = phd.c;
            const int total = cp.ne0 * cp.nrows;
            for (int e = blockIdx.x * (DS_NCW * 32) + threadIdx.x; e < total; e += gridDim.x * DS_NCW * 32) {
                if (cp.add) cp.dst[e] = __fadd_rn(ldcg_f((const float *)cp.src + e), ldcg_f(cp.src2 + e));
                else {
                    const int r = e / cp.ne0, i = e % cp.ne0;
                    cp.dst[e] = ldcg_f((const float *)(cp.src + (size_t)__ldcg(cp.idx + r) * cp.nb1) + i);
                }
            }
        }
        DSPROF(2);
        if (ph + 1 < P.nphases) grid_sync(P, nbar);
        else named_bar_sync(1, DS_NCW * 32);
        DSPROF(3);
    }
#undef DSPROF
    // ---- leave the barrier counter at zero for the next launch: the last CTA to finish resets it ----
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&P.sync[1], 1u) == gridDim.x - 1) { P.sync[0] = 0; P.sync[1] = 0; __threadfence(); }
    }
}

// ---------------------------------------------------------------------------------------------- host side
struct DsProgramImpl {
    std::vector<uint8_t> key;            // the DsPhase array (host copy): programs are cached by content
    DsPhase *dev = nullptr;
    DsParams params = {};
    int grid = 0, mask = 0, nphases = 0;
    size_t smem = 0;
};

}  // namespace

struct DsProgram : DsProgramImpl {};

struct DsCache {
    std::vector<DsProgram *> progs;
    unsigned int *sync = nullptr;
};

static DsCache *ds_cache(b200_ctx *ctx) {
    if (!ctx->dstep_cache) {
        DsCache *c = new DsCache();
        if (cudaMalloc((void **)&c->sync, 64) != cudaSuccess) { cudaGetLastError(); delete c; return nullptr; }
        cudaMemset(c->sync, 0, 64);
        ctx->dstep_cache = c;
    }
    return (DsCache *)ctx->dstep_cache;
}

void dstep_cache_free(b200_ctx *ctx) {
    DsCache *c = (DsCache *)ctx->dstep_cache;
    if (!c) return;
    for (DsProgram *p : c->progs) { if (p->dev) cudaFree(p->dev); delete p; }
    if (c->sync) cudaFree(c->sync);
    delete c;
    ctx->dstep_cache = nullptr;
}

static int type_bit(int t) { return t == B200_TYPE_Q4_K ? TB_Q4_K : t == B200_TYPE_Q5_K ? TB_Q5_K : t == B200_TYPE_Q6_K ? TB_Q6_K : 0; }

bool dstep_gemv_eligible(const b200_ctx *ctx, const GemvSegDesc *segs, int nseg, int64_t K, const GemvActDesc &ga, int ncols) {
    if (ncols != 1 || nseg < 1 || nseg > GEMV_MAX_SEG || K <= 0 || (K & 255) || K > 32768) return false;
    if (ga.mode != ACT_F32 && ga.mode != ACT_F32_NORM && ga.mode != ACT_F32_SWIGLU) return false;
    if (((uintptr_t)ga.x & 15) || (ga.mode != ACT_F32 && ((uintptr_t)ga.x2 & 15))) return false;
    for (int s = 0; s < nseg; s++) {
        const GemvSegDesc &g = segs[s];
        if (!type_bit(g.type) || g.expert_id || g.dbgP || g.N < ctx->sm_count || g.N > (1 << 22)) return false;
        if (g.type != B200_TYPE_Q6_K && (((uintptr_t)g.W & 15) || (g.rb & 15))) return false;
        if (g.type == B200_TYPE_Q6_K && ((uintptr_t)g.W & 1)) return false;
    }
    return true;
}

bool dstep_attn_eligible(const b200_ctx *ctx, const RopeStoreDesc &rs, const b200_op &fa) {
    (void)ctx;
    const b200_tensor &q = fa.src[0], &k = fa.src[1], &v = fa.src[2], &m = fa.src[3];
    if (rs.T != 1 || rs.D != 128 || q.ne[0] != 128 || q.ne[1] != 1) return false;
    if (rs.kv_type != B200_TYPE_F16 && rs.kv_type != B200_TYPE_Q8_0 && rs.kv_type != B200_TYPE_Q4_0) return false;
    if (k.type != rs.kv_type || v.type != rs.kv_type) return false;
    const int H = rs.H, Hkv = rs.Hkv;
    if (Hkv <= 0 || H % Hkv) return false;
    const int gq = H / Hkv;
    if (gq != 1 && gq != 2 && gq != 4 && gq != 8) return false;
    if (q.ne[2] != H || k.ne[2] != Hkv || v.ne[2] != Hkv || k.ne[1] != v.ne[1]) return false;
    if (fa.n_src < 4 || !m.data || m.type != B200_TYPE_F16) return false;
    float max_bias, softcap;
    memcpy(&max_bias, &fa.params[1], 4); memcpy(&softcap, &fa.params[2], 4);
    if (max_bias != 0.0f || softcap != 0.0f) return false;
    if ((rs.q_out != (float *)q.data)) return false;                       // the roped q feeds this flash_attn and nothing else (matcher)
    if (rs.kv_type == B200_TYPE_F16 && ((k.nb[1] & 7) || (v.nb[1] & 7) || (k.nb[2] & 7) || (v.nb[2] & 7) || ((uintptr_t)k.data & 7) || ((uintptr_t)v.data & 7))) return false;
    if (rs.kv_type != B200_TYPE_F16 && ((k.nb[1] & 1) || (v.nb[1] & 1) || (k.nb[2] & 1) || (v.nb[2] & 1))) return false;
    // the new token's cache rows must be cells of these cache views
    if (k.nb[1] == 0 || v.nb[1] == 0) return false;
    return true;
}

bool dstep_copy_eligible(const b200_op &op) {
    const b200_tensor &s = op.src[0], &i = op.src[1], &d = op.dst;
    if (op.op == B200_OP_ADD)         // dense same-shape f32 add (the residual add the matcher could not fold into a GEMV)
        return s.type == B200_TYPE_F32 && i.type == B200_TYPE_F32 && d.type == B200_TYPE_F32 && tensor_is_contiguous(s) && tensor_is_contiguous(i) &&
               tensor_is_contiguous(d) && tensor_nelements(s) == tensor_nelements(d) && tensor_nelements(i) == tensor_nelements(d) &&
               !memcmp(s.ne, d.ne, sizeof(s.ne)) && !memcmp(i.ne, d.ne, sizeof(i.ne)) && tensor_nelements(d) < (1 << 20);
    return op.op == B200_OP_GET_ROWS && s.type == B200_TYPE_F32 && d.type == B200_TYPE_F32 && i.type == B200_TYPE_I32 && s.nb[0] == 4 &&
           tensor_is_contiguous(d) && d.ne[0] == s.ne[0] && d.ne[2] == 1 && d.ne[3] == 1 && i.ne[0] == d.ne[1] && i.ne[1] == 1 && s.ne[2] == 1 &&
           s.ne[3] == 1 && i.nb[0] == 4 && d.ne[0] * d.ne[1] < (1 << 24);
}

template <int TYPES>
static int ds_launch_t(b200_ctx *ctx, DsProgram *pr) {
    auto kern = b200_decode_step_kernel<TYPES>;
    static bool attr_set[16] = {false};
    if (!attr_set[ctx->device & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        attr_set[ctx->device & 15] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)pr->grid);
    cfg.blockDim = dim3(DS_THREADS);
    cfg.dynamicSmemBytes = pr->smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;          // all CTAs co-resident, or the launch fails (never a deadlocked barrier)
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, pr->params));
    ctx->launches++;
    return B200_OK;
}

int dstep_launch(b200_ctx *ctx, DsProgram *pr) {
    pr->params.prof = (unsigned long long *)ctx->prof_buf;
    switch (pr->mask) {
        case TB_Q4_K: case TB_Q4_K | TB_Q6_K: case TB_Q6_K: return ds_launch_t<TB_Q4_K | TB_Q6_K>(ctx, pr);
        case TB_Q5_K: case TB_Q5_K | TB_Q6_K: return ds_launch_t<TB_Q5_K | TB_Q6_K>(ctx, pr);
        default: return ds_launch_t<TB_Q4_K | TB_Q5_K | TB_Q6_K>(ctx, pr);
    }
}

int dstep_prepare(b200_ctx *ctx, const std::vector<DsNode> &nodes, DsProgram **out) {
    DsCache *cache = ds_cache(ctx);
    if (!cache) return B200_ERR_ALLOC;
    const int G = ctx->sm_count;
    // ---- shared-memory plan: barriers | s_red | part | phase descriptor | activation area | ring ----
    // The activation area holds the quantised input vector of ONE phase in the padded layouts its decoders read (aq64: 272 B per
    // 256-block for q4_K / q5_K, aq128: 144 B per 128 for q6_K), sized for the most demanding phase; the attention phase reuses it.
    int mask = 0;
    uint32_t act_bytes = (8 + 2) * 128 * 4 + 31 * 132 * 4;           // attention: [gq + 2][128] f32 + [<= 31][132] f32
    for (const DsNode &n : nodes) {
        if (n.kind != DS_GEMV) continue;
        bool n64 = false, n128 = false;
        for (int s = 0; s < n.nseg; s++) { mask |= type_bit(n.seg[s].type); if (n.seg[s].type == B200_TYPE_Q6_K) n128 = true; else n64 = true; }
        const uint32_t nb = (uint32_t)(n.K >> 8);
        uint32_t o = 0;
        if (n64) o += (nb * 272 + 15) & ~15u;
        if (n128) o += (2 * nb * 144 + 15) & ~15u;
        o += (nb * 4 + 15) & ~15u;
        if (n64) o += (nb * 16 + 15) & ~15u;
        if (n128) o += (nb * 32 + 15) & ~15u;
        act_bytes = std::max(act_bytes, o);
    }
    DsParams P = {};
    uint32_t off = 64 * 8 + 32 * 8;
    P.off_part = off; off += DS_MAX_PIECES * DS_PART_ROWS * 4;
    P.off_desc = off; off += DS_DESC_BYTES + 64;
    P.off_pgeo = off; off += 32 * 64;
    off = (off + 127) & ~127u;
    P.off_act = off; off += act_bytes;
    off = (off + 127) & ~127u;
    P.off_ring = off;
    // leave >= 12 KB of the SM's 228 KB to the L1 (the x / residual / KV reads of the consumers)
    const size_t budget = std::min((size_t)ctx->smem_optin, (size_t)214 * 1024);
    if ((size_t)off + 8 * 4352 > budget) { b200_set_error("dstep: no room for the ring (activation area %u bytes)", act_bytes); return B200_ERR_UNSUPPORTED; }
    const size_t ring_bytes = budget - off;
    uint32_t stage = (uint32_t)(ring_bytes / DS_MAX_STAGES) & ~127u;
    if (stage < 4352) stage = 4352;
    int ns = (int)(ring_bytes / stage);
    if (ns > DS_MAX_STAGES) ns = DS_MAX_STAGES;
    P.nstages = ns; P.stage_bytes = stage;
    // ---- phases ----
    std::vector<DsPhase> prog;
    for (const DsNode &n : nodes) {
        DsPhase ph;
        memset(&ph, 0, sizeof(ph));
        ph.kind = n.kind;
        if (n.kind == DS_GEMV) {
            DsGemv &g = ph.g;
            g.nseg = n.nseg; g.K = (int)n.K; g.act_mode = n.act.mode; g.eps = n.act.eps; g.x = n.act.x; g.x2 = n.act.x2;
            const int nb = (int)(n.K >> 8);
            int split_rows = 0;
            bool n64 = false, n128 = false;
            for (int s = 0; s < n.nseg; s++) { if (n.seg[s].type == B200_TYPE_Q6_K) n128 = true; else n64 = true; }
            {
                uint32_t o = 0;
                g.off_aq64 = g.off_aq128 = g.off_s32 = g.off_s16 = 0xffffffffu;
                if (n64) { g.off_aq64 = o; o += ((uint32_t)nb * 272 + 15) & ~15u; }
                if (n128) { g.off_aq128 = o; o += (2 * (uint32_t)nb * 144 + 15) & ~15u; }
                g.off_ad = o; o += ((uint32_t)nb * 4 + 15) & ~15u;
                if (n64) { g.off_s32 = o; o += ((uint32_t)nb * 16 + 15) & ~15u; }
                if (n128) { g.off_s16 = o; o += ((uint32_t)nb * 32 + 15) & ~15u; }
            }
            for (int s = 0; s < n.nseg; s++) {
                DsSeg &d = g.seg[s];
                const GemvSegDesc &sd = n.seg[s];
                d.W = sd.W; d.dst = sd.dst; d.residual = sd.residual; d.rb = (uint32_t)sd.rb; d.type = sd.type; d.N = (int)sd.N;
                const uint32_t bbytes = sd.type == B200_TYPE_Q4_K ? 144u : sd.type == B200_TYPE_Q5_K ? 176u : 210u;
                d.S = 1; d.nbp = nb; d.R = 1; d.lgR = 0;
                if (d.rb + 32 <= stage) {
                    int R = (int)((stage - 32) / d.rb);
                    R = R > 8 ? 8 : R;
                    if (sd.type != B200_TYPE_Q6_K && R >= 2 && nb > 16) R = 1;          // pair decoder walks 16 blocks per half-warp pass: keep long rows on 32 lanes
                    while ((2 << d.lgR) <= R) d.lgR++;
                    d.R = 1 << d.lgR;
                } else {
                    while (d.S < DS_MAX_PIECES && (uint32_t)((nb + d.S - 1) / d.S) * bbytes + 32 > stage) d.S++;
                    d.nbp = (nb + d.S - 1) / d.S;
                    if ((uint32_t)d.nbp * bbytes + 32 > stage) { b200_set_error("dstep: row of %u bytes does not fit %d pieces", d.rb, d.S); return B200_ERR_UNSUPPORTED; }
                    split_rows += (int)(sd.N / G) + 1;
                }
                d.q = (int)(sd.N / G); d.rem = (int)(sd.N % G);
            }
            if (split_rows > DS_PART_ROWS) { b200_set_error("dstep: %d K-split rows per CTA", split_rows); return B200_ERR_UNSUPPORTED; }
            prog.push_back(ph);
        } else if (n.kind == DS_ATTN) {
            DsAttn &a = ph.a;
            const b200_tensor &k = n.fa.src[1], &v = n.fa.src[2], &m = n.fa.src[3];
            a.q = n.rs.q; a.k = n.rs.k; a.v = n.rs.v; a.pos = n.rs.pos; a.ff = n.rs.freq_factors;
            a.kc = (const char *)k.data; a.vc = (const char *)v.data; a.k_nb1 = k.nb[1]; a.k_nb2 = k.nb[2]; a.v_nb1 = v.nb[1]; a.v_nb2 = v.nb[2];
            a.mask = (const char *)m.data;
            a.k_dst = n.rs.k_dst; a.v_dst = n.rs.v_dst; a.k_dst_ind = n.rs.k_dst_ind; a.v_dst_ind = n.rs.v_dst_ind;
            a.out = (float *)n.fa.dst.data;
            a.H = n.rs.H; a.Hkv = n.rs.Hkv; a.gq = a.H / a.Hkv; a.n_kv = (int)k.ne[1];
            a.kvt = n.rs.kv_type == B200_TYPE_F16 ? KVT_F16 : n.rs.kv_type == B200_TYPE_Q8_0 ? KVT_Q8_0 : KVT_Q4_0;
            a.npw = DS_NCW / a.gq;
            int nsplit = G / a.Hkv;
            const int max_split = (a.n_kv + 31) / 32;
            if (nsplit > max_split) nsplit = max_split;
            if (nsplit < 1) nsplit = 1;
            a.nsplit = nsplit;
            a.len = ((a.n_kv + nsplit - 1) / nsplit + 31) / 32 * 32;
            a.nsplit = (a.n_kv + a.len - 1) / a.len;
            memcpy(&a.scale, &n.fa.params[0], 4);
            a.rp = make_rope_params(n.rs.rope_params);
            a.part = (float *)ctx->get_scratch(SCRATCH_FATTN, (size_t)(3 * ctx->sm_count + 64) * 16 * (128 + 2) * 4);
            if (!a.part) return B200_ERR_ALLOC;
            prog.push_back(ph);
            DsPhase cph = ph;
            cph.kind = DS_COMBINE;
            prog.push_back(cph);
        } else {
            DsCopy &cp = ph.c;
            cp.add = n.cp.op == B200_OP_ADD;
            cp.src = (const char *)n.cp.src[0].data; cp.nb1 = n.cp.src[0].nb[1];
            if (cp.add) { cp.src2 = (const float *)n.cp.src[1].data; cp.ne0 = (int)tensor_nelements(n.cp.dst); cp.nrows = 1; }
            else { cp.idx = (const int32_t *)n.cp.src[1].data; cp.ne0 = (int)n.cp.dst.ne[0]; cp.nrows = (int)n.cp.dst.ne[1]; }
            cp.dst = (float *)n.cp.dst.data;
            prog.push_back(ph);
        }
    }
    // ---- cached by content ----
    const size_t nbytes = prog.size() * sizeof(DsPhase);
    for (DsProgram *p : cache->progs)
        if (p->key.size() == nbytes && p->params.nstages == P.nstages && p->params.off_ring == P.off_ring && !memcmp(p->key.data(), prog.data(), nbytes)) { *out = p; return B200_OK; }
    if (ctx->capturing) { b200_set_error("dstep: program upload during graph capture"); return B200_ERR_FAILED; }
    // the kernel stages whole descriptors with 16-byte loads: pad the device array by one descriptor
    if (cache->progs.size() >= 256) {            // bounded: captured graphs may reference programs, so drop those first
        void graph_cache_free(b200_ctx *);
        cudaStreamSynchronize(ctx->stream);
        graph_cache_free(ctx);
        for (DsProgram *p : cache->progs) { if (p->dev) cudaFree(p->dev); delete p; }
        cache->progs.clear();
    }
    DsProgram *pr = new DsProgram();
    pr->key.assign((const uint8_t *)prog.data(), (const uint8_t *)prog.data() + nbytes);
    if (cudaMalloc((void **)&pr->dev, nbytes + DS_DESC_BYTES) != cudaSuccess) { cudaGetLastError(); delete pr; b200_set_error("dstep: program alloc"); return B200_ERR_ALLOC; }
    CUDA_TRY(cudaMemcpyAsync(pr->dev, pr->key.data(), nbytes, cudaMemcpyHostToDevice, ctx->stream));
    P.prog = pr->dev; P.nphases = (int)prog.size(); P.sync = cache->sync;
    pr->params = P; pr->grid = G; pr->mask = mask; pr->nphases = (int)prog.size();
    pr->smem = (size_t)P.off_ring + (size_t)P.nstages * P.stage_bytes;
    cache->progs.push_back(pr);
    *out = pr;
    return B200_OK;
}
