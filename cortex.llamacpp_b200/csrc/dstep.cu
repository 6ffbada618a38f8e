// dstep.cu -- the persistent decode-step kernel: ONE launch per decoded token, no kernel boundaries and no grid barriers.
//
// Round 1 ran a batch-1 llama step as 226 launches (6 fused kernels per layer).  The streaming GEMV itself reached 99% of the
// HBM copy peak on large matrices, but a step averaged 0.51: every kernel boundary drains the memory pipeline (HBM idles
// while the next launch starts, reloads its activations and re-quantises them), and 4096 x 4096 matrices are only 1.4 us of
// HBM time (profiles/r1_gemv_diag.md).  This kernel removes the boundaries instead of shaving them:
//   * one CTA per SM (grid = SM count, cooperative launch), 31 consumer warps + 1 producer warp, alive for the whole step;
//   * the step is a PROGRAM of phases in device memory (DsPhase[]: GEMV over 1-3 weight matrices with rms_norm / silu*up fused
//     around it and the residual in its epilogue; rope + KV store + split-KV attention; split merge; row gather / add), built by
//     graph.cu's matcher from the ggml op list -- the same fused nodes the per-launch kernels take;
//   * weights are constants, so the producer warp streams them through ONE shared-memory ring (cp.async.bulk + mbarrier,
//     L2 evict_first) for the whole program without ever waiting for a phase to finish: while the consumers wait for their
//     inputs or run attention, the ring (31 stages, ~170 KB per SM = 25 MB chip-wide, ~4 us of HBM time) keeps filling with the
//     NEXT matmul's rows.  HBM streams across every dependency of the step;
//   * phases hand their results to each other through FLAGGED MAILBOXES instead of barriers: every value that crosses CTAs is
//     one aligned 8-byte word {payload, epoch} written with a single store and polled with L2 loads (the scheme of NCCL's LL
//     protocol).  A reader that sees the epoch sees the payload: no fence, no atomic, no barrier -- a measured grid barrier on
//     148 SMs costs 1.1 us even without memory ordering and 1.5-2 us with it (tools/micro/barrier_bench.cu), an L2 hand-off
//     ~0.4 us.  Mailbox slots rotate over 4 phases, epochs are unique per (launch, phase);
//   * activations are quantised ONCE: the CTAs that own a 256-block of the next matmul's input (16-56 of the 148) poll the f32
//     mailbox, apply rms_norm*w (the matmul's producer already folded silu(gate)*up into what it published), quantise
//     exactly like quantize_row_q8_K_ref and publish the q8_K block; every CTA then polls ~10-34 KB of q8 words into its
//     shared-memory layouts instead of re-reading 16-114 KB of f32 and re-quantising it 148 times;
//   * long rows (K = 14336: 8-12 KB) are cut into K-pieces of one stage each so that all 31 consumer warps stay busy; rows are
//     collected in shared memory and a row's pieces are added in a fixed order at the end of the phase (bit-reproducible);
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
constexpr int DS_CT = DS_NCW * 32;         // consumer threads
constexpr int DS_THREADS = (DS_NCW + 1) * 32;
constexpr int DS_MAX_STAGES = 31;          // one producer lane per stage, stage s is consumed by warp s
constexpr int DS_MAX_PIECES = 8;
constexpr int DS_SLOTS = 4;                // mailbox slots (phase p uses slot p % 4)
constexpr int DS_QBLK = 80;                // 8-byte words per published q8_K block: 64 quant words, d, 4 words of per-32 sums, 8 words of per-16 sums, pad
constexpr int TB_Q4_K = 1, TB_Q5_K = 2, TB_Q6_K = 4;
enum { KVT_F16 = 0, KVT_Q8_0 = 1, KVT_Q4_0 = 2 };
enum { OUT_NONE = 0, OUT_ROWS = 1, OUT_SWIGLU = 2 };
typedef unsigned long long u64;

struct DsSeg {
    const uint8_t *W;
    float *        dst;
    const float *  residual;
    uint32_t       rb;                     // bytes per row
    int            type, N;
    int            R, lgR;                 // rows per chunk (power of two; 1 when the row is K-split)
    int            S, nbp;                 // K-pieces per row and 256-blocks per piece
    int            q, rem;                 // N = q * grid + rem: CTA c owns rows [c*q + min(c, rem), ...)
    uint32_t       mf_off;                 // OUT_ROWS: element offset of this segment's rows inside the f32 mailbox slot
};
struct DsGemv {
    int            nseg, K, act_mode;      // act_mode: what the quantiser applies (ACT_F32 / ACT_F32_NORM / ACT_F32_SWIGLU)
    float          eps;
    const float *  x, *x2;                 // plain inputs (in_flag == 0), or the norm weight in x2
    int            in_flag, in_ph;         // 1: x (and x2 for swiglu) are read from the f32 mailbox written by phase in_ph
    uint32_t       in_off, in_off2;        // element offsets inside that mailbox slot
    int            out_mode, pstride;      // how rows are published; row-buffer plane stride (this CTA's rows of the phase, padded)
    uint32_t       off_aq64, off_aq128, off_ad, off_s32, off_s16;     // activation layouts of THIS phase inside the activation area (~0u = not needed)
    DsSeg          seg[GEMV_MAX_SEG];
};
struct DsAttn {
    const float *  q, *k, *v;              // plain raw q/k/v of the token (in_flag == 0)
    int            in_flag, in_ph;         // 1: read from the f32 mailbox of phase in_ph: q at in_off, k at in_off + H*D, v after k
    uint32_t       in_off;
    const int32_t *pos;
    const float *  ff;                     // rope frequency factors (optional)
    const char *   kc, *vc;                // cache views of the FLASH_ATTN_EXT op: cell stride nb1, head stride nb2
    uint64_t       k_nb1, k_nb2, v_nb1, v_nb2;
    const char *   mask;                   // f16 [n_kv] row of the token
    void *         k_dst, *v_dst;          // cache rows of the new token ([Hkv][D] in the cache type) ...
    void *const *  k_dst_ind, *const *v_dst_ind;      // ... or where to read them from (CUDA-graph replay, graph.cu)
    float *        out;                    // attention output [H * D] (plain copy)
    int            H, Hkv, gq, n_kv, kvt, nsplit, len, npw;
    float          scale;
    RopeParams     rp;
};
struct DsCopy { const char *src; const int32_t *idx; const float *src2; uint64_t nb1; float *dst; int ne0, nrows, add; };   // get_rows (f32 rows), or dst = src + src2
struct alignas(16) DsPhase {
    int kind, sync_before;                 // sync_before: this phase reads plain memory written by other CTAs earlier in the launch -> grid barrier first
    union { DsGemv g; DsAttn a; DsCopy c; };
};

struct DsParams {
    const DsPhase *prog;
    int            nphases;
    unsigned int * sync;                   // [0] barrier arrivals, [1] exits, [2] error flag, [3] launch sequence number (epoch base)
    u64 *          mf;                     // f32 mailboxes   [DS_SLOTS][mf_slot]  words {f32 bits, epoch}
    u64 *          mq;                     // q8_K mailboxes  [DS_SLOTS][mq_slot]  words {4 payload bytes, epoch}, DS_QBLK words per 256-block
    u64 *          mp;                     // attention split partials [Hkv * nsplit][gq][132] words {f32 bits, epoch}
    uint32_t       mf_slot, mq_slot;
    int            nstages, max_inflight;
    uint32_t       stage_bytes;
    uint32_t       off_rowbuf, off_desc, off_pgeo, off_act, off_ring;
    unsigned long long *prof;              // debug: [grid][nphases][8] %globaltimer stamps: 0 phase start, 1 inputs in shared memory, 2 work done (thread 0), 3 phase end, 4-7 attention steps
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

// ---- flagged mailbox words: one aligned 8-byte store / load carries payload and epoch together (single-copy atomic) ----
__device__ __forceinline__ u64 ld_word(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_word(u64 *p, uint32_t payload, uint32_t ep) {
    const u64 v = ((u64)ep << 32) | payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __noinline__ u64 poll_slow(const u64 *p, uint32_t ep, unsigned *err) {
    const long long t0 = clock64();
    u64 v;
    unsigned ns = 32;
    do {
        __nanosleep(ns);                      // back off: tens of thousands of threads wait on the L2 at phase boundaries
        if (ns < 256) ns <<= 1;
        v = ld_word(p);
        if (clock64() - t0 > (1ll << 32)) { atomicExch(err, 1u); break; }        // ~2 s: the producer is missing; give up with garbage, the host sees the flag
    } while ((uint32_t)(v >> 32) != ep);
    return v;
}
// one lane of the warp waits for a sentinel word, then the whole warp proceeds (its own words are still verified afterwards)
__device__ __forceinline__ void warp_wait_word(const u64 *p, uint32_t ep, unsigned *err, int lane) {
    if (lane == 0) { const u64 v = ld_word(p); if ((uint32_t)(v >> 32) != ep) poll_slow(p, ep, err); }
    __syncwarp();
}
__device__ __forceinline__ uint32_t poll_word(const u64 *p, uint32_t ep, unsigned *err) {
    u64 v = ld_word(p);
    if ((uint32_t)(v >> 32) != ep) v = poll_slow(p, ep, err);
    return (uint32_t)v;
}

// per-CTA geometry of a GEMV phase (shared memory): rows [lo, hi) of every segment, first chunk index, row-buffer offset
struct PhaseGeo { int lo[GEMV_MAX_SEG], hi[GEMV_MAX_SEG], ch0[GEMV_MAX_SEG + 1], poff[GEMV_MAX_SEG], pad[3]; };
static_assert(sizeof(PhaseGeo) == 64, "PhaseGeo");
__device__ __forceinline__ int phase_geo(const DsGemv &g, int c, PhaseGeo *o) {       // returns the chunk count
    int po = 0, ch = 0;
    o->ch0[0] = 0;
#pragma unroll
    for (int s = 0; s < GEMV_MAX_SEG; s++) {
        int lo = 0, hi = 0, n = 0;
        const int po_s = po;
        if (s < g.nseg) {
            const int q = g.seg[s].q, rem = g.seg[s].rem;
            lo = c * q + min(c, rem);
            hi = (c + 1) * q + min(c + 1, rem);
            const int rows = hi - lo;
            n = ((rows + g.seg[s].R - 1) >> g.seg[s].lgR) * g.seg[s].S;
            po += rows;
        }
        ch += n;
        o->lo[s] = lo; o->hi[s] = hi; o->poff[s] = po_s; o->ch0[s + 1] = ch;
    }
    return ch;
}

// grid-wide barrier (consumers only).  Only used around the rare phases that read plain memory written by other CTAs (row
// gathers / adds of llama.cpp's last layer): everything else goes through the mailboxes.  1.5-2 us on 148 SMs.
__device__ __forceinline__ void grid_sync(const DsParams &P, int &nbar) {
    named_bar_sync(1, DS_CT);
    nbar++;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&P.sync[0], 1u);
        const unsigned target = (unsigned)nbar * gridDim.x;
        const long long t0 = clock64();
        while (ld_acquire_u32(&P.sync[0]) < target) {
            if (clock64() - t0 > (1ll << 32)) { atomicExch(&P.sync[2], 1u); break; }
        }
        __threadfence();
    }
    named_bar_sync(1, DS_CT);
}

// ---------------------------------------------------------------------------------------------- activation quantiser (block owners)
__device__ __forceinline__ void apply_mode(int mode, float (&v)[8], const float (&w)[8], float norm_scale) {
    if (mode == ACT_F32_NORM) {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __fmul_rn(__fmul_rn(v[j], norm_scale), w[j]);
    } else if (mode == ACT_F32_SWIGLU) {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __fmul_rn(ggml_silu_lane(v[j]), w[j]);
    }
}
// 8 consecutive f32 of a phase input: plain memory, or mailbox words of epoch ep_in
__device__ __forceinline__ void load8(const DsParams &P, const float *plain, const u64 *mail, uint32_t ep_in, int e0, float (&v)[8]) {
    if (mail) {
        u64 r[8];
        const long long t0 = clock64();
        unsigned ns = 32;
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 8; j++) r[j] = ld_word(mail + e0 + j);       // all eight in flight, then check the epochs
#pragma unroll
            for (int j = 0; j < 8; j++) ok &= (uint32_t)(r[j] >> 32) == ep_in;
            if (__all_sync(0xffffffffu, ok)) break;
            if (clock64() - t0 > (1ll << 32)) { atomicExch(&P.sync[2], 1u); break; }
            __nanosleep(ns);
            if (ns < 256) ns <<= 1;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __uint_as_float((uint32_t)r[j]);
    } else {
        const float4 a = __ldcg((const float4 *)(plain + e0)), b = __ldcg((const float4 *)(plain + e0 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
}
// one warp: q8_K of 256 values (lane owns v[0..7]) -> block b of this CTA's shared-memory layouts
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
// Phase input -> q8_K in shared memory, bit-exact vs quantize_row_q8_K_ref, after rms_norm(x)*w or silu(g)*u when the phase asks
// for it.  The input is read straight from the producing phase's mailbox (one L2 round trip once the words carry the epoch): a
// dedicated quantiser hop was measured slower -- under the streaming load every dependent L2 access costs ~1 us, so the number of
// hops on the critical path matters more than the 148-fold re-quantisation (profiles/r2_dstep_timeline.md).
__device__ __forceinline__ void gemv_prologue(const DsParams &P, const ActOff &A, const DsGemv &g, uint32_t ep_in, uint8_t *smem, double *s_red, int warp, int lane) {
    const int nchunk = g.K >> 8, mode = g.act_mode;
    const u64 *mail = g.in_flag ? P.mf + (size_t)(g.in_ph % DS_SLOTS) * P.mf_slot + g.in_off : nullptr;
    const u64 *mail2 = g.in_flag ? P.mf + (size_t)(g.in_ph % DS_SLOTS) * P.mf_slot + g.in_off2 : nullptr;
    auto norm_scale = [&](double s) -> float {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) s_red[warp] = s;
        named_bar_sync(2, DS_CT);
        double t = lane < DS_NCW ? s_red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        named_bar_sync(2, DS_CT);                 // s_red is reused by the next phase
        return __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn((float)(t / (double)g.K), g.eps)));
    };
    auto weights = [&](int e0, float (&w)[8]) {
        if (mode == ACT_F32_NORM) {
            const float4 a = *(const float4 *)(g.x2 + e0), b = *(const float4 *)(g.x2 + e0 + 4);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        } else if (mode == ACT_F32_SWIGLU) load8(P, g.x2, mail2, ep_in, e0, w);
    };
    if (nchunk <= 2 * DS_NCW) {
        // every llama shape at batch 1: a warp keeps its <= 2 blocks in registers across the rms_norm reduction (x is read once)
        float va[8], vb[8], wa[8], wb[8];
        const int ba = warp, bb = warp + DS_NCW;
#pragma unroll
        for (int j = 0; j < 8; j++) va[j] = vb[j] = wa[j] = wb[j] = 0.0f;
        if (mail && mode != ACT_F32_SWIGLU) {
            // both blocks' mailbox words in flight together: one L2 round trip once the producers have published
            u64 ra[8], rb[8];
            const u64 *pa = mail + ba * 256 + lane * 8, *pb = mail + bb * 256 + lane * 8;
            const long long t0 = clock64();
            unsigned nsl = 32;
            for (;;) {
                bool ok = true;
#pragma unroll
                for (int j = 0; j < 8; j++) { ra[j] = ba < nchunk ? ld_word(pa + j) : ((u64)ep_in << 32); rb[j] = bb < nchunk ? ld_word(pb + j) : ((u64)ep_in << 32); }
#pragma unroll
                for (int j = 0; j < 8; j++) ok &= (uint32_t)(ra[j] >> 32) == ep_in && (uint32_t)(rb[j] >> 32) == ep_in;
                if (__all_sync(0xffffffffu, ok)) break;
                if (clock64() - t0 > (1ll << 32)) { atomicExch(&P.sync[2], 1u); break; }
                __nanosleep(nsl);
                if (nsl < 256) nsl <<= 1;
            }
#pragma unroll
            for (int j = 0; j < 8; j++) { va[j] = __uint_as_float((uint32_t)ra[j]); vb[j] = __uint_as_float((uint32_t)rb[j]); }
            if (ba < nchunk) weights(ba * 256 + lane * 8, wa);
            if (bb < nchunk) weights(bb * 256 + lane * 8, wb);
        } else {
            if (ba < nchunk) { load8(P, g.x, mail, ep_in, ba * 256 + lane * 8, va); weights(ba * 256 + lane * 8, wa); }
            if (bb < nchunk) { load8(P, g.x, mail, ep_in, bb * 256 + lane * 8, vb); weights(bb * 256 + lane * 8, wb); }
        }
        float ns = 1.0f;
        if (mode == ACT_F32_NORM) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < 8; j++) s += (double)__fmul_rn(va[j], va[j]);
#pragma unroll
            for (int j = 0; j < 8; j++) s += (double)__fmul_rn(vb[j], vb[j]);
            ns = norm_scale(s);
        }
        if (ba < nchunk) { apply_mode(mode, va, wa, ns); quant_block_to_smem(A, smem, ba, lane, va); }
        if (bb < nchunk) { apply_mode(mode, vb, wb, ns); quant_block_to_smem(A, smem, bb, lane, vb); }
        return;
    }
    float ns = 1.0f;
    if (mode == ACT_F32_NORM) {
        double s = 0.0;
        for (int b = warp; b < nchunk; b += DS_NCW) {
            float v[8];
            load8(P, g.x, mail, ep_in, b * 256 + lane * 8, v);
#pragma unroll
            for (int j = 0; j < 8; j++) s += (double)__fmul_rn(v[j], v[j]);
        }
        ns = norm_scale(s);
    }
#pragma unroll 1
    for (int b = warp; b < nchunk; b += DS_NCW) {
        float v[8], w[8];
        load8(P, g.x, mail, ep_in, b * 256 + lane * 8, v);
        weights(b * 256 + lane * 8, w);
        apply_mode(mode, v, w, ns);
        quant_block_to_smem(A, smem, b, lane, v);
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
        const float x0 = src[ia], x1 = src[ib];
        dst[ia] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
        dst[ib] = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
    } else {
        dst[i0] = src[i0];
        dst[i0 + 1] = src[i0 + 1];
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

// ---- attention walk: one warp = one query head of the group x one interleave of the split's cells; lane owns 4 of the 128 dims.
// Cells are handled in batches of NB whose mask / K / V loads are all in flight together (every dependent L2 access costs ~1 us
// under the streaming load): at depth 512 a warp's cells are one batch, i.e. one round trip.
constexpr int DS_NB = 4;
struct CellBatch { float mv[DS_NB]; unsigned live; uint32_t kx[DS_NB], ky[DS_NB], vx[DS_NB], vy[DS_NB]; };      // live: bit j = cell j takes part
struct WalkState { float q0, q1, q2, q3, dq; int qi; float o0, o1, o2, o3, mrow, lrow; };

template <int KVT>
__device__ __forceinline__ void load_batch(const DsAttn &a, int hk, int cb, int step, int c1, int skip_cell, int lane, CellBatch &B) {
    const char *kb = a.kc + (size_t)hk * a.k_nb2, *vb = a.vc + (size_t)hk * a.v_nb2;
    const int blk = lane >> 3, e0 = (lane & 7) * 4;        // quantised rows: 32-element block and offset inside it
    constexpr int BB = KVT == KVT_Q8_0 ? 34 : 18;
    B.live = 0;
#pragma unroll
    for (int j = 0; j < DS_NB; j++) {
        const int c = cb + j * step;
        const bool in = c < c1 && c != skip_cell;
        B.mv[j] = 0.0f;
        B.kx[j] = B.ky[j] = B.vx[j] = B.vy[j] = 0;
        if (in) {
            // mask, K and V requested together: a masked cell costs two wasted 256-byte reads, a dependent load would cost ~1 us
            const __half mh = __ldcg((const __half *)(a.mask + (size_t)c * 2));
            const uint8_t *kr = (const uint8_t *)(kb + (size_t)c * a.k_nb1), *vr = (const uint8_t *)(vb + (size_t)c * a.v_nb1);
            if (KVT == KVT_F16) {
                const uint2 kk = __ldcg((const uint2 *)(kr + lane * 8)), vv = __ldcg((const uint2 *)(vr + lane * 8));
                B.kx[j] = kk.x; B.ky[j] = kk.y; B.vx[j] = vv.x; B.vy[j] = vv.y;
            } else {
                const uint8_t *kblk = kr + blk * BB, *vblk = vr + blk * BB;
                B.ky[j] = __ldcg((const unsigned short *)kblk); B.vy[j] = __ldcg((const unsigned short *)vblk);
                const int o = KVT == KVT_Q8_0 ? 2 + e0 : 2 + (e0 & 15);
                B.kx[j] = ld_u16x2_cg(kblk + o); B.vx[j] = ld_u16x2_cg(vblk + o);
            }
            B.mv[j] = __half2float(mh);
            if (!(__hisinf(mh) && B.mv[j] < 0.0f)) B.live |= 1u << j;
        }
    }
}
template <int KVT>
__device__ __forceinline__ void walk_init(const float *qrow, int lane, WalkState &W) {
    const float4 qf = *(const float4 *)(qrow + lane * 4);
    W.q0 = qf.x; W.q1 = qf.y; W.q2 = qf.z; W.q3 = qf.w; W.dq = 0.0f; W.qi = 0;
    if (KVT == KVT_F16) {        // the CPU rounds Q to f16 for an f16 K (vec_dot_type)
        W.q0 = __half2float(__float2half_rn(W.q0)); W.q1 = __half2float(__float2half_rn(W.q1));
        W.q2 = __half2float(__float2half_rn(W.q2)); W.q3 = __half2float(__float2half_rn(W.q3));
    } else {                     // ... and quantises it to q8_0 for a quantised K: block = 8 lanes x 4 elements
        float amax = fmaxf(fmaxf(fabsf(W.q0), fabsf(W.q1)), fmaxf(fabsf(W.q2), fabsf(W.q3)));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
        const float d = __fdiv_rn(amax, 127.0f), id = amax != 0.0f ? __fdiv_rn(127.0f, amax) : 0.0f;
        W.dq = __half2float(__float2half_rn(d));
        const int a0 = __float2int_rn(__fmul_rn(W.q0, id)), a1 = __float2int_rn(__fmul_rn(W.q1, id));
        const int a2 = __float2int_rn(__fmul_rn(W.q2, id)), a3 = __float2int_rn(__fmul_rn(W.q3, id));
        W.qi = (a0 & 0xff) | ((a1 & 0xff) << 8) | ((a2 & 0xff) << 16) | ((a3 & 0xff) << 24);
    }
    W.o0 = W.o1 = W.o2 = W.o3 = 0.0f; W.mrow = -INFINITY; W.lrow = 0.0f;
}
template <int KVT>
__device__ __forceinline__ void consume_batch(const DsAttn &a, int lane, const CellBatch &B, WalkState &W) {
    const int e0 = (lane & 7) * 4;
#pragma unroll
    for (int j = 0; j < DS_NB; j++) {
        if ((B.live >> j) & 1u) {                   // uniform across the warp
        float s, v0, v1, v2, v3;
        if (KVT == KVT_F16) {
            const float2 k01 = __half22float2(*(const __half2 *)&B.kx[j]), k23 = __half22float2(*(const __half2 *)&B.ky[j]);
            const float2 v01 = __half22float2(*(const __half2 *)&B.vx[j]), v23 = __half22float2(*(const __half2 *)&B.vy[j]);
            s = fmaf(W.q3, k23.y, fmaf(W.q2, k23.x, fmaf(W.q1, k01.y, __fmul_rn(W.q0, k01.x))));
            s = warp_reduce_sum(s);
            v0 = v01.x; v1 = v01.y; v2 = v23.x; v3 = v23.y;
        } else {
            const float dk = __half2float(__ushort_as_half((unsigned short)B.ky[j])), dv = __half2float(__ushort_as_half((unsigned short)B.vy[j]));
            uint32_t kw = B.kx[j], vw = B.vx[j];
            if (KVT == KVT_Q4_0) {
                kw = __vsub4(e0 < 16 ? (kw & 0x0f0f0f0fu) : ((kw >> 4) & 0x0f0f0f0fu), 0x08080808u);
                vw = __vsub4(e0 < 16 ? (vw & 0x0f0f0f0fu) : ((vw >> 4) & 0x0f0f0f0fu), 0x08080808u);
            }
            int isum = __dp4a((int)kw, W.qi, 0);
            isum += __shfl_xor_sync(0xffffffffu, isum, 1);
            isum += __shfl_xor_sync(0xffffffffu, isum, 2);
            isum += __shfl_xor_sync(0xffffffffu, isum, 4);
            s = __fmul_rn((float)isum, __fmul_rn(dk, W.dq));                 // every lane of a block holds the block's term
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            v0 = __fmul_rn((float)(int8_t)(vw & 0xff), dv); v1 = __fmul_rn((float)(int8_t)((vw >> 8) & 0xff), dv);
            v2 = __fmul_rn((float)(int8_t)((vw >> 16) & 0xff), dv); v3 = __fmul_rn((float)(int8_t)(vw >> 24), dv);
        }
        s = fmaf(s, a.scale, B.mv[j]);
        const float mnew = fmaxf(W.mrow, s);
        const float corr = W.mrow == -INFINITY ? 0.0f : expf(W.mrow - mnew), pw = expf(s - mnew);
        W.mrow = mnew;
        W.lrow = fmaf(W.lrow, corr, pw);
        W.o0 = fmaf(W.o0, corr, pw * v0); W.o1 = fmaf(W.o1, corr, pw * v1); W.o2 = fmaf(W.o2, corr, pw * v2); W.o3 = fmaf(W.o3, corr, pw * v3);
        }
    }
}

// one (kv head, split) unit: rope, KV store (owner split), online-softmax walk over the cells, merge of the CTA's warps
template <int KVT>
__device__ __forceinline__ void attn_unit_t(const DsParams &P, const DsAttn &a, uint32_t ep_in, uint32_t ep, uint8_t *scr, int warp, int lane, int ph) {
#define APROF(slot) do { if (P.prof && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); P.prof[((size_t)blockIdx.x * P.nphases + ph) * 8 + (slot)] = t_; } } while (0)
    constexpr int D = 128;
    constexpr int MB = D + 4;                         // per-warp partial rows stay 16-byte aligned
    const int unit = blockIdx.x;
    if (unit >= a.Hkv * a.nsplit) return;
    const int hk = unit / a.nsplit, sp = unit % a.nsplit;
    const int c0 = sp * a.len, c1 = min(a.n_kv, c0 + a.len);
    float *sq = (float *)scr;                         // [gq][D] roped q
    float *skv = sq + a.gq * D;                       // [2][D]: roped k, v of the new token (owner split only)
    float *sraw = skv + 2 * D;                        // [gq + 2][D] raw q (group), k, v of this kv head
    float *mbuf = sraw + (a.gq + 2) * D;              // [gq * npw][MB] per-warp partials
    const int tid = threadIdx.x;
    char *kd = (char *)(a.k_dst_ind ? *a.k_dst_ind : a.k_dst), *vd = (char *)(a.v_dst_ind ? *a.v_dst_ind : a.v_dst);
    const int cell = (int)((kd - a.kc) / (long long)a.k_nb1);
    const bool owner = cell >= c0 && cell < c1;
    const int hg = warp % a.gq, sub = warp / a.gq;
    const bool walker = sub < a.npw;
    const int p = __ldcg(a.pos);
    // ---- raw q (gq heads), k, v of this kv head -> shared memory (mailbox of the qkv phase, or plain) ----
    {
        const u64 *mail = a.in_flag ? P.mf + (size_t)(a.in_ph % DS_SLOTS) * P.mf_slot + a.in_off : nullptr;
        const int nq = a.gq * D;
        for (int t = tid; t < nq + (owner ? 2 * D : 0); t += DS_CT) {
            size_t e; const float *pl;
            if (t < nq) { e = (size_t)hk * nq + t; pl = a.q + e; }
            else if (t < nq + D) { e = (size_t)a.H * D + (size_t)hk * D + (t - nq); pl = a.k + (size_t)hk * D + (t - nq); }
            else { e = (size_t)(a.H + a.Hkv) * D + (size_t)hk * D + (t - nq - D); pl = a.v + (size_t)hk * D + (t - nq - D); }
            sraw[t] = mail ? __uint_as_float(poll_word(mail + e, ep_in, &P.sync[2])) : ldcg_f(pl);
        }
    }
    named_bar_sync(3, DS_CT);
    APROF(4);
    // ---- rope: gq query heads (+ the k head), one pair per thread ----
    for (int t = tid; t < (a.gq + 1) * (D / 2); t += DS_CT) {
        const int h = t / (D / 2), ip = t % (D / 2);
        if (h < a.gq) rope_pair(a.rp, a.ff, p, ip, sraw + h * D, sq + h * D);
        else if (owner) rope_pair(a.rp, a.ff, p, ip, sraw + a.gq * D, skv);
    }
    if (owner) for (int t = tid; t < D; t += DS_CT) skv[D + t] = sraw[(a.gq + 1) * D + t];
    named_bar_sync(3, DS_CT);
    if (owner) {
        const size_t rowb = KVT == KVT_F16 ? (size_t)D * 2 : KVT == KVT_Q8_0 ? (size_t)(D / 32) * 34 : (size_t)(D / 32) * 18;
        if (tid < D) store_kv_row(KVT, skv, kd + (size_t)hk * rowb, tid);
        else if (tid < 2 * D) store_kv_row(KVT, skv + D, vd + (size_t)hk * rowb, tid - D);
    }
    if (owner) { __threadfence_block(); named_bar_sync(3, DS_CT); }       // the new cell is in the cache (this CTA reads it back through L2)
    APROF(5);
    if (walker) {
        WalkState W;
        CellBatch B;
        walk_init<KVT>(sq + hg * D, lane, W);
#pragma unroll 1
        for (int cb = c0 + sub; cb < c1; cb += DS_NB * a.npw) {
            load_batch<KVT>(a, hk, cb, a.npw, c1, -1, lane, B);
            consume_batch<KVT>(a, lane, B, W);
        }
        float *mb = mbuf + (size_t)(hg * a.npw + sub) * MB;
        *(float4 *)(mb + lane * 4) = make_float4(W.o0, W.o1, W.o2, W.o3);
        if (lane == 0) { mb[D] = W.mrow; mb[D + 1] = W.lrow; }
    }
    APROF(6);
    named_bar_sync(3, DS_CT);
    APROF(7);
    // ---- merge the interleaves of every head (fixed order) and publish the split partial ----
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
        u64 *pp = P.mp + ((size_t)unit * a.gq + warp) * MB;
        st_word(pp + lane * 4, __float_as_uint(r0), ep); st_word(pp + lane * 4 + 1, __float_as_uint(r1), ep);
        st_word(pp + lane * 4 + 2, __float_as_uint(r2), ep); st_word(pp + lane * 4 + 3, __float_as_uint(r3), ep);
        if (lane == 0) { st_word(pp + D, __float_as_uint(M), ep); st_word(pp + D + 1, __float_as_uint(L), ep); }
    }
}
#undef APROF
__device__ __forceinline__ void attn_unit(const DsParams &P, const DsAttn &a, uint32_t ep_in, uint32_t ep, uint8_t *scr, int warp, int lane, int ph) {
    if (a.kvt == KVT_F16) attn_unit_t<KVT_F16>(P, a, ep_in, ep, scr, warp, lane, ph);
    else if (a.kvt == KVT_Q8_0) attn_unit_t<KVT_Q8_0>(P, a, ep_in, ep, scr, warp, lane, ph);
    else attn_unit_t<KVT_Q4_0>(P, a, ep_in, ep, scr, warp, lane, ph);
}

// merge of the KV splits of one head: CTA h < H, warps 0..3 take 32 dims each; publishes the attention output.
// All words of all splits are requested at once (nsplit <= 18 independent loads per lane), checked, re-read only if stale.
__device__ __forceinline__ void attn_combine(const DsParams &P, const DsAttn &a, uint32_t ep_in, uint32_t ep, int ph, int warp, int lane) {
    constexpr int D = 128;
    constexpr int MB = D + 4;
    const int h = blockIdx.x;
    if (h >= a.H || warp >= 4) return;
    const int hk = h / a.gq, hg = h % a.gq;
    const u64 *base = P.mp + ((size_t)(hk * a.nsplit) * a.gq + hg) * MB;
    const size_t sstride = (size_t)a.gq * MB;
    const int d = warp * 32 + lane;
    // lane s holds (M_s, L_s); every lane then takes its dim of the splits, 12 at a time (all 12 loads in flight together)
    u64 wm = 0, wl = 0;
    {
        const long long t0 = clock64();
        unsigned nsl = 32;
        for (;;) {
            bool ok = true;
            if (lane < a.nsplit) { wm = ld_word(base + lane * sstride + D); wl = ld_word(base + lane * sstride + D + 1); ok = (uint32_t)(wm >> 32) == ep_in && (uint32_t)(wl >> 32) == ep_in; }
            if (__all_sync(0xffffffffu, ok)) break;
            if (clock64() - t0 > (1ll << 32)) { atomicExch(&P.sync[2], 1u); break; }
            __nanosleep(nsl);
            if (nsl < 256) nsl <<= 1;
        }
    }
    const float ms_l = lane < a.nsplit ? __uint_as_float((uint32_t)wm) : -INFINITY, ls_l = lane < a.nsplit ? __uint_as_float((uint32_t)wl) : 0.0f;
    const float M = warp_reduce_max(ms_l);
    float L = 0.0f, acc = 0.0f;
    constexpr int PS = 12;
#pragma unroll 1
    for (int s0 = 0; s0 < a.nsplit; s0 += PS) {
        u64 wa[PS];
#pragma unroll
        for (int s = 0; s < PS; s++) wa[s] = s0 + s < a.nsplit ? ld_word(base + (s0 + s) * sstride + d) : ((u64)ep_in << 32);
#pragma unroll
        for (int s = 0; s < PS; s++) {
            if (s0 + s < a.nsplit) {
                // (M_s, L_s) carried the epoch, so the split has been published; a row word still in flight is re-read
                if ((uint32_t)(wa[s] >> 32) != ep_in) wa[s] = poll_slow(base + (s0 + s) * sstride + d, ep_in, &P.sync[2]);
                const float ms = __shfl_sync(0xffffffffu, ms_l, s0 + s), ls = __shfl_sync(0xffffffffu, ls_l, s0 + s);
                const float f = ms == -INFINITY ? 0.0f : expf(ms - M);
                L = fmaf(ls, f, L);
                acc = fmaf(__uint_as_float((uint32_t)wa[s]), f, acc);
            }
        }
    }
    const float r = acc / L;
    a.out[(size_t)h * D + d] = r;
    st_word(P.mf + (size_t)(ph % DS_SLOTS) * P.mf_slot + (size_t)h * D + d, __float_as_uint(r), ep);
}

// ---------------------------------------------------------------------------------------------- the kernel
constexpr int DS_DESC_BYTES = 288;
static_assert(sizeof(DsPhase) == DS_DESC_BYTES, "phase descriptor size");

template <int TYPES>
__global__ void __launch_bounds__(DS_THREADS, 1) b200_decode_step_kernel(const DsParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full  = (uint64_t *)smem;                      // [32]
    uint64_t *empty = full + 32;                             // [32]
    double *  s_red = (double *)(empty + 32);                // [32]
    float *   rowbuf = (float *)(smem + P.off_rowbuf);       // [pieces][pstride] dot products of this CTA's rows
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
            bool want = false, landed = true;
            if (active) {
                while (gi >= ce) {
                    ph++;
                    if (ph >= P.nphases) { active = false; break; }
                    if (P.prog[ph].kind == DS_GEMV) { cb = ce; ce = cb + phase_geo(P.prog[ph].g, c, geo); }
                }
                landed = use == 0 || mbar_test_wait(&full[lane], (use - 1) & 1);
                want = active && (use == 0 || mbar_test_wait(&empty[lane], (use - 1) & 1));
            }
            // Little's law caps what is worth having in flight: bytes in flight = bandwidth x latency.  Re-arming all 31 stages at
            // once (170 KB per SM, 25 MB chip-wide) only lengthens the memory system's queues to ~4 us -- for these copies AND for
            // every latency-critical mailbox / residual / KV load of the consumers.  So at most `max_inflight` stages are
            // outstanding per SM (enough to saturate HBM); the rest of the ring is buffering capacity, not queue depth.
            const int inflight = __popc(__ballot_sync(0xffffffffu, !landed));
            const unsigned wants = __ballot_sync(0xffffffffu, want);
            const int rank = __popc(wants & ((1u << lane) - 1u));
            if (want && inflight + rank < P.max_inflight) {
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
        return;
    }

    // ====================================================================== consumers
    asm volatile("bar.sync 4, %0;" ::"r"(DS_THREADS) : "memory");        // mbarriers are initialised
    int nbar = 0, use = 0;
    int cbm = 0;                                                          // (global chunk index where the current phase starts) mod ns
    const DsPhase &phd = *(const DsPhase *)(smem + P.off_desc);           // the current phase's descriptor, staged in shared memory
    PhaseGeo *sgeo = (PhaseGeo *)(smem + P.off_desc + DS_DESC_BYTES);     // ... and this CTA's row / chunk geometry of it
    const uint32_t ep_base = __ldcg(&P.sync[3]) * 4096u + 1u;             // epoch of phase ph = ep_base + ph: unique per (launch, phase)
#define DSPROF(slot) do { if (P.prof && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); P.prof[((size_t)blockIdx.x * P.nphases + ph) * 8 + (slot)] = t_; } } while (0)
    for (int ph = 0; ph < P.nphases; ph++) {
        DSPROF(0);
        // every consumer reads the descriptor dozens of times per chunk: one copy per phase instead of L2 round trips
        if (threadIdx.x < DS_DESC_BYTES / 16) ((uint4 *)(smem + P.off_desc))[threadIdx.x] = __ldg((const uint4 *)&P.prog[ph] + threadIdx.x);
        named_bar_sync(1, DS_CT);
        if (phd.sync_before) grid_sync(P, nbar);
        const uint32_t ep = ep_base + (uint32_t)ph;
        if (phd.kind == DS_GEMV) {
            const DsGemv &g = phd.g;
            if (threadIdx.x == 32) phase_geo(g, c, sgeo);
            const uint32_t ep_in = ep_base + (uint32_t)g.in_ph;
            ActOff A;
            A.n64 = g.off_aq64 != 0xffffffffu; A.n128 = g.off_aq128 != 0xffffffffu;
            A.aq64 = P.off_act + g.off_aq64; A.aq128 = P.off_act + g.off_aq128; A.ad = P.off_act + g.off_ad; A.s32 = P.off_act + g.off_s32; A.s16 = P.off_act + g.off_s16;
            gemv_prologue(P, A, g, ep_in, smem, s_red, warp, lane);
            named_bar_sync(1, DS_CT);
            DSPROF(1);
            const int nch = sgeo->ch0[GEMV_MAX_SEG], nseg = g.nseg, pstride = g.pstride;
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
                float *rb_out = rowbuf + pc * pstride + sgeo->poff[sidx] - sgeo->lo[sidx];       // rowbuf[piece][this CTA's row index]
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
                        if (bl == 0 && mine) rb_out[row0 + r + sub] = acc;
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
                        if (lane == 0) rb_out[row0 + r] = acc;
                    }
                }
                use++;
            }
            named_bar_sync(1, DS_CT);
            DSPROF(2);
            // ---- epilogue: pieces in piece order, residual, plain store, mailbox word for the next phase ----
            u64 *mf = P.mf + (size_t)(ph % DS_SLOTS) * P.mf_slot;
            const int out_mode = g.out_mode;
#pragma unroll
            for (int s = 0; s < GEMV_MAX_SEG; s++) {
                if (s < nseg) {
                    const int S = g.seg[s].S, lo = sgeo->lo[s], rows = sgeo->hi[s] - sgeo->lo[s], po = sgeo->poff[s];
                    const float *residual = g.seg[s].residual;
                    float *dst = g.seg[s].dst;
                    const uint32_t mo = g.seg[s].mf_off;
                    for (int r = threadIdx.x; r < rows; r += DS_CT) {
                        float v = rowbuf[po + r];
                        for (int pc = 1; pc < S; pc++) v = __fadd_rn(v, rowbuf[pc * pstride + po + r]);
                        if (residual) v = __fadd_rn(v, ldcg_f(residual + lo + r));
                        dst[lo + r] = v;
                        if (out_mode == OUT_ROWS) st_word(mf + mo + lo + r, __float_as_uint(v), ep);
                        else if (out_mode == OUT_SWIGLU) {
                            if (s == 0) rowbuf[po + r] = v;                  // gate, kept for the fold below (same thread reads it back)
                            else st_word(mf + lo + r, __float_as_uint(__fmul_rn(ggml_silu_lane(rowbuf[sgeo->poff[0] + r]), v)), ep);
                        }
                    }
                }
            }
            cbm = (cbm + nch) % ns;
        } else if (phd.kind == DS_ATTN) {
            attn_unit(P, phd.a, ep_base + (uint32_t)phd.a.in_ph, ep, smem + P.off_act, warp, lane, ph);
            DSPROF(2);
        } else if (phd.kind == DS_COMBINE) {
            attn_combine(P, phd.a, ep_base + (uint32_t)(ph - 1), ep, ph, warp, lane);
            DSPROF(2);
        } else {        // DS_COPY: dst[r][:] = src[idx[r]][:] (f32 rows), or dst = src + src2; plain memory, bracketed by grid barriers
            const DsCopy &cp = phd.c;
            const int total = cp.ne0 * cp.nrows;
            for (int e = blockIdx.x * DS_CT + threadIdx.x; e < total; e += gridDim.x * DS_CT) {
                if (cp.add) cp.dst[e] = __fadd_rn(ldcg_f((const float *)cp.src + e), ldcg_f(cp.src2 + e));
                else {
                    const int r = e / cp.ne0, i = e % cp.ne0;
                    cp.dst[e] = ldcg_f((const float *)(cp.src + (size_t)__ldcg(cp.idx + r) * cp.nb1) + i);
                }
            }
            DSPROF(2);
        }
        named_bar_sync(1, DS_CT);           // the descriptor / activation area / row buffer are reused by the next phase
        DSPROF(3);
    }
#undef DSPROF
    // ---- last CTA out: barrier counter back to zero, next launch gets a new epoch base ----
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&P.sync[1], 1u) == gridDim.x - 1) { P.sync[0] = 0; P.sync[1] = 0; P.sync[3] = P.sync[3] + 1; __threadfence(); }
    }
}

// ---------------------------------------------------------------------------------------------- host side
constexpr uint32_t DS_MF_SLOT = 131072;        // f32 mailbox words per slot (largest published vector: gate|up rows of a 70B model = 57344)
constexpr uint32_t DS_MQ_BLOCKS = 160;         // q8 blocks per slot (K <= 40960)
constexpr uint32_t DS_MP_WORDS = 148 * 8 * 132;

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
    u64 *mf = nullptr, *mq = nullptr, *mp = nullptr;
};

static DsCache *ds_cache(b200_ctx *ctx) {
    if (!ctx->dstep_cache) {
        DsCache *c = new DsCache();
        const size_t nmf = (size_t)DS_SLOTS * DS_MF_SLOT * 8, nmq = (size_t)DS_SLOTS * DS_MQ_BLOCKS * DS_QBLK * 8, nmp = (size_t)DS_MP_WORDS * 8;
        if (cudaMalloc((void **)&c->sync, 64) != cudaSuccess || cudaMalloc((void **)&c->mf, nmf) != cudaSuccess ||
            cudaMalloc((void **)&c->mq, nmq) != cudaSuccess || cudaMalloc((void **)&c->mp, nmp) != cudaSuccess) {
            cudaGetLastError();
            if (c->sync) cudaFree(c->sync); if (c->mf) cudaFree(c->mf); if (c->mq) cudaFree(c->mq); if (c->mp) cudaFree(c->mp);
            delete c;
            return nullptr;
        }
        cudaMemset(c->sync, 0, 64); cudaMemset(c->mf, 0, nmf); cudaMemset(c->mq, 0, nmq); cudaMemset(c->mp, 0, nmp);      // epoch 0 = never written
        ctx->dstep_cache = c;
    }
    return (DsCache *)ctx->dstep_cache;
}

void dstep_cache_free(b200_ctx *ctx) {
    DsCache *c = (DsCache *)ctx->dstep_cache;
    if (!c) return;
    for (DsProgram *p : c->progs) { if (p->dev) cudaFree(p->dev); delete p; }
    if (c->sync) cudaFree(c->sync);
    if (c->mf) cudaFree(c->mf);
    if (c->mq) cudaFree(c->mq);
    if (c->mp) cudaFree(c->mp);
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

// a value produced earlier in the program: [lo, hi) bytes of plain memory, by which phase, how it was published
struct DsProduced { uintptr_t lo, hi; int ph, kind, N, seg; uint32_t mf_off; bool mailed; };

int dstep_prepare(b200_ctx *ctx, const std::vector<DsNode> &nodes, DsProgram **out) {
    DsCache *cache = ds_cache(ctx);
    if (!cache) return B200_ERR_ALLOC;
    const int G = ctx->sm_count;
    // ---- shared-memory plan: barriers | s_red | row buffer | phase descriptor + geometry | activation area | ring ----
    // The activation area holds the quantised input vector of ONE phase in the padded layouts its decoders read (aq64: 272 B per
    // 256-block for q4_K / q5_K, aq128: 144 B per 128 for q6_K), sized for the most demanding phase; the attention phase reuses it.
    int mask = 0;
    uint32_t act_bytes = (2 * 8 + 4) * 128 * 4 + 31 * 132 * 4;           // attention: q, raw q/k/v, roped k/v + [<= 31][132] f32
    for (const DsNode &n : nodes) {
        if (n.kind != DS_GEMV) continue;
        bool n64 = false, n128 = false;
        for (int s = 0; s < n.nseg; s++) { mask |= type_bit(n.seg[s].type); if (n.seg[s].type == B200_TYPE_Q6_K) n128 = true; else n64 = true; }
        const uint32_t nb = (uint32_t)(n.K >> 8);
        if (nb > DS_MQ_BLOCKS) { b200_set_error("dstep: K=%lld", (long long)n.K); return B200_ERR_UNSUPPORTED; }
        uint32_t o = 0;
        if (n64) o += (nb * 272 + 15) & ~15u;
        if (n128) o += (2 * nb * 144 + 15) & ~15u;
        o += (nb * 4 + 15) & ~15u;
        if (n64) o += (nb * 16 + 15) & ~15u;
        if (n128) o += (nb * 32 + 15) & ~15u;
        act_bytes = std::max(act_bytes, o);
    }
    // ring geometry first (the row buffer depends on how rows are cut into pieces): assume a 6 KB row buffer, fix up below
    DsParams P = {};
    auto plan = [&](uint32_t rowbuf_bytes, uint32_t &stage, int &ns) -> bool {
        uint32_t off = 64 * 8 + 32 * 8;
        P.off_rowbuf = off; off += (rowbuf_bytes + 15) & ~15u;
        P.off_desc = off; off += DS_DESC_BYTES + 64;
        P.off_pgeo = off; off += 32 * 64;
        off = (off + 127) & ~127u;
        P.off_act = off; off += act_bytes;
        off = (off + 127) & ~127u;
        P.off_ring = off;
        // leave ~12 KB of the SM's 228 KB to the L1 (residual / KV / mailbox reads of the consumers)
        const size_t budget = std::min((size_t)ctx->smem_optin, (size_t)214 * 1024);
        if ((size_t)off + 8 * 3456 > budget) return false;
        const size_t ring_bytes = budget - off;
        stage = (uint32_t)(ring_bytes / DS_MAX_STAGES) & ~127u;
        if (stage < 3456) stage = 3456;                              // >= 16 q6_K blocks
        ns = (int)(ring_bytes / stage);
        if (ns > DS_MAX_STAGES) ns = DS_MAX_STAGES;
        return true;
    };
    uint32_t stage = 0, rowbuf_bytes = 6144;
    int ns = 0;
    std::vector<DsPhase> prog;
    for (int attempt = 0; attempt < 2; attempt++) {
        if (!plan(rowbuf_bytes, stage, ns)) { b200_set_error("dstep: no room for the ring (activation area %u bytes)", act_bytes); return B200_ERR_UNSUPPORTED; }
        prog.clear();
        std::vector<DsProduced> made;
        uint32_t rowbuf_need = 0;
        auto producer_of = [&](const void *p) -> const DsProduced * {
            const DsProduced *best = nullptr;
            for (const DsProduced &d : made) if ((uintptr_t)p >= d.lo && (uintptr_t)p < d.hi) best = &d;      // the latest writer wins
            return best;
        };
        bool sync_next = false;          // the previous phase was a plain-memory phase: the next one starts with a grid barrier
        for (const DsNode &n : nodes) {
            DsPhase ph;
            memset(&ph, 0, sizeof(ph));
            ph.kind = n.kind;
            ph.sync_before = sync_next ? 1 : 0;
            sync_next = false;
            const int pi = (int)prog.size();
            if (n.kind == DS_GEMV) {
                DsGemv &g = ph.g;
                g.nseg = n.nseg; g.K = (int)n.K; g.act_mode = n.act.mode; g.eps = n.act.eps; g.x = n.act.x; g.x2 = n.act.x2;
                const int nb = (int)(n.K >> 8);
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
                // ---- where does the input come from?  mailbox of the previous phase, or plain memory (+ grid barrier if another CTA wrote it) ----
                const DsProduced *px = producer_of(n.act.x);
                const DsProduced *px2 = n.act.mode == ACT_F32_SWIGLU ? producer_of(n.act.x2) : nullptr;
                if (n.act.mode == ACT_F32_SWIGLU && px && px2 && px->ph == pi - 1 && px2->ph == pi - 1 && px->kind == DS_GEMV && px->seg == 0 && px2->seg == 1 &&
                    px->N == px2->N && px->N == n.K && (uintptr_t)n.act.x == px->lo && (uintptr_t)n.act.x2 == px2->lo && prog[pi - 1].g.nseg == 2) {
                    prog[pi - 1].g.out_mode = OUT_SWIGLU;          // the producer publishes silu(gate) * up; this phase quantises it as is
                    g.act_mode = ACT_F32; g.in_flag = 1; g.in_ph = pi - 1; g.in_off = 0;
                } else if (n.act.mode != ACT_F32_SWIGLU && px && px->ph == pi - 1 && px->mailed && (uintptr_t)n.act.x == px->lo && px->N >= n.K) {
                    g.in_flag = 1; g.in_ph = pi - 1; g.in_off = px->mf_off;
                } else if (px || px2) ph.sync_before = 1;          // produced in this launch but not through the mailbox: plain read after a grid barrier
                int rows_cta = 0, maxS = 1;
                uint32_t mo = 0;
                for (int s = 0; s < n.nseg; s++) {
                    DsSeg &d = g.seg[s];
                    const GemvSegDesc &sd = n.seg[s];
                    d.W = sd.W; d.dst = sd.dst; d.residual = sd.residual; d.rb = (uint32_t)sd.rb; d.type = sd.type; d.N = (int)sd.N;
                    d.mf_off = mo; mo += (uint32_t)sd.N;
                    const uint32_t bbytes = sd.type == B200_TYPE_Q4_K ? 144u : sd.type == B200_TYPE_Q5_K ? 176u : 210u;
                    d.S = 1; d.nbp = nb; d.R = 1; d.lgR = 0;
                    if (d.rb + 32 <= stage) {
                        int R = (int)((stage - 32) / d.rb);
                        R = R > 8 ? 8 : R;
                        if (sd.type != B200_TYPE_Q6_K && R >= 2 && nb > 16) R = 1;          // pair decoder walks 16 blocks per half-warp pass: keep long rows on 32 lanes
                        while ((2 << d.lgR) <= R) d.lgR++;
                        d.R = 1 << d.lgR;
                    } else {
                        // K-pieces: q4_K / q5_K decode a block per lane (<= 32 blocks per piece, balanced); q6_K decodes 128 weights per lane
                        // (pieces of 16 blocks = 32 items, so every pass uses all lanes)
                        int nbp_max = (int)((stage - 32) / bbytes);
                        if (sd.type == B200_TYPE_Q6_K) nbp_max = nbp_max / 16 * 16; else if (nbp_max > 32) nbp_max = 32;
                        if (nbp_max < 1) { b200_set_error("dstep: stage of %u bytes too small", stage); return B200_ERR_UNSUPPORTED; }
                        d.S = (nb + nbp_max - 1) / nbp_max;
                        d.nbp = sd.type == B200_TYPE_Q6_K ? nbp_max : (nb + d.S - 1) / d.S;
                        if (d.S > DS_MAX_PIECES) { b200_set_error("dstep: row of %u bytes needs %d pieces", d.rb, d.S); return B200_ERR_UNSUPPORTED; }
                    }
                    d.q = (int)(sd.N / G); d.rem = (int)(sd.N % G);
                    rows_cta += d.q + 1;
                    maxS = std::max(maxS, d.S);
                    // the residual is read plainly by the CTA that owns the row: fine if the same row partition wrote it (same N, a GEMV
                    // of this launch) or nobody in this launch did; otherwise it needs the grid barrier
                    if (sd.residual) { const DsProduced *pr = producer_of(sd.residual); if (pr && !(pr->kind == DS_GEMV && pr->N == d.N && (uintptr_t)sd.residual == pr->lo)) ph.sync_before = 1; }
                }
                if (mo > DS_MF_SLOT) { b200_set_error("dstep: %u rows in one phase", mo); return B200_ERR_UNSUPPORTED; }
                g.pstride = (rows_cta + 3) & ~3;
                rowbuf_need = std::max(rowbuf_need, (uint32_t)g.pstride * (uint32_t)maxS * 4u);
                g.out_mode = OUT_ROWS;
                for (int s = 0; s < n.nseg; s++)
                    made.push_back(DsProduced{(uintptr_t)n.seg[s].dst, (uintptr_t)n.seg[s].dst + (size_t)n.seg[s].N * 4, pi, DS_GEMV, (int)n.seg[s].N, s, g.seg[s].mf_off, true});
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
                if (nsplit > 24) nsplit = 24;                      // attn_combine keeps one word per split in registers
                const int max_split = (a.n_kv + 31) / 32;
                if (nsplit > max_split) nsplit = max_split;
                if (nsplit < 1) nsplit = 1;
                a.len = (a.n_kv + nsplit - 1) / nsplit;
                a.nsplit = (a.n_kv + a.len - 1) / a.len;
                if ((size_t)a.Hkv * a.nsplit * a.gq * 132 > DS_MP_WORDS) { b200_set_error("dstep: attention partials"); return B200_ERR_UNSUPPORTED; }
                memcpy(&a.scale, &n.fa.params[0], 4);
                a.rp = make_rope_params(n.rs.rope_params);
                // q, k, v: the three segments of the previous (qkv) phase, in that order?
                const DsProduced *pq = producer_of(n.rs.q), *pk = producer_of(n.rs.k), *pv = producer_of(n.rs.v);
                if (pq && pk && pv && pq->ph == pi - 1 && pk->ph == pi - 1 && pv->ph == pi - 1 && pq->mailed && pq->seg == 0 && pk->seg == 1 && pv->seg == 2 &&
                    (uintptr_t)n.rs.q == pq->lo && (uintptr_t)n.rs.k == pk->lo && (uintptr_t)n.rs.v == pv->lo && pq->N == a.H * 128 && pk->N == a.Hkv * 128 && pv->N == a.Hkv * 128) {
                    a.in_flag = 1; a.in_ph = pi - 1; a.in_off = 0;
                } else if (pq || pk || pv) ph.sync_before = 1;
                prog.push_back(ph);
                DsPhase cph = ph;
                cph.kind = DS_COMBINE; cph.sync_before = 0;
                made.push_back(DsProduced{(uintptr_t)a.out, (uintptr_t)a.out + (size_t)a.H * 128 * 4, pi + 1, DS_COMBINE, a.H * 128, 0, 0, true});
                prog.push_back(cph);
            } else {
                DsCopy &cp = ph.c;
                cp.add = n.cp.op == B200_OP_ADD;
                cp.src = (const char *)n.cp.src[0].data; cp.nb1 = n.cp.src[0].nb[1];
                if (cp.add) { cp.src2 = (const float *)n.cp.src[1].data; cp.ne0 = (int)tensor_nelements(n.cp.dst); cp.nrows = 1; }
                else { cp.idx = (const int32_t *)n.cp.src[1].data; cp.ne0 = (int)n.cp.dst.ne[0]; cp.nrows = (int)n.cp.dst.ne[1]; }
                cp.dst = (float *)n.cp.dst.data;
                ph.sync_before = 1;                    // plain reads of what other CTAs wrote ...
                sync_next = true;                      // ... and plain writes the next phase may read
                made.push_back(DsProduced{(uintptr_t)cp.dst, (uintptr_t)cp.dst + (size_t)cp.ne0 * cp.nrows * 4, pi, DS_COPY, cp.ne0 * cp.nrows, 0, 0, false});
                prog.push_back(ph);
            }
        }
        if (!prog.empty() && prog.back().kind == DS_GEMV) prog.back().g.out_mode = OUT_NONE;      // nobody in this launch reads the last matmul
        if (rowbuf_need <= rowbuf_bytes) break;
        rowbuf_bytes = rowbuf_need;                      // (the output matmul of a large vocabulary: ~870 rows per CTA) plan again
    }
    P.nstages = ns; P.stage_bytes = stage;
    static const int env_inflight = getenv("GGML_B200_DS_INFLIGHT") ? atoi(getenv("GGML_B200_DS_INFLIGHT")) : 0;
    P.max_inflight = env_inflight > 0 ? env_inflight : ns;          // measured: capping the outstanding copies below the ring depth only costs bandwidth (profiles/r2_dstep_timeline.md)
    P.mf = cache->mf; P.mq = cache->mq; P.mp = cache->mp; P.mf_slot = DS_MF_SLOT; P.mq_slot = DS_MQ_BLOCKS * DS_QBLK;
    // ---- cached by content ----
    const size_t nbytes = prog.size() * sizeof(DsPhase);
    for (DsProgram *p : cache->progs)
        if (p->key.size() == nbytes && p->params.nstages == P.nstages && p->params.off_ring == P.off_ring && p->params.stage_bytes == P.stage_bytes &&
            !memcmp(p->key.data(), prog.data(), nbytes)) { *out = p; return B200_OK; }
    if (ctx->capturing) { b200_set_error("dstep: program upload during graph capture"); return B200_ERR_FAILED; }
    if (cache->progs.size() >= 256) {            // bounded: captured graphs may reference programs, so drop those first
        void graph_cache_free(b200_ctx *);
        cudaStreamSynchronize(ctx->stream);
        graph_cache_free(ctx);
        for (DsProgram *p : cache->progs) { if (p->dev) cudaFree(p->dev); delete p; }
        cache->progs.clear();
    }
    DsProgram *pr = new DsProgram();
    pr->key.assign((const uint8_t *)prog.data(), (const uint8_t *)prog.data() + nbytes);
    // the kernel stages whole descriptors with 16-byte loads: pad the device array by one descriptor
    if (cudaMalloc((void **)&pr->dev, nbytes + DS_DESC_BYTES) != cudaSuccess) { cudaGetLastError(); delete pr; b200_set_error("dstep: program alloc"); return B200_ERR_ALLOC; }
    CUDA_TRY(cudaMemcpyAsync(pr->dev, pr->key.data(), nbytes, cudaMemcpyHostToDevice, ctx->stream));
    P.prog = pr->dev; P.nphases = (int)prog.size(); P.sync = cache->sync;
    pr->params = P; pr->grid = G; pr->mask = mask; pr->nphases = (int)prog.size();
    pr->smem = (size_t)P.off_ring + (size_t)P.nstages * P.stage_bytes;
    cache->progs.push_back(pr);
    *out = pr;
    return B200_OK;
}
