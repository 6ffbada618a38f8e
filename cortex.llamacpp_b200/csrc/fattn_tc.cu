// fattn_tc.cu -- FLASH_ATTN_EXT for prompt-sized query blocks on the 5th-generation tensor cores (tcgen05.mma kind::f16, TMEM).
//
// Replaces flash_attn_ext_f16 (ggml-cuda/fattn-mma-f16.cuh:804-936: mma.sync tiles) for ubatches of >= 64 query columns at head
// size 128 over f16 / q8_0 / q4_0 caches, with the CPU's operand semantics (ggml_compute_forward_flash_attn_ext_f16,
// ggml-cpu.c:12270-12420) kept exactly:
//   * one CTA = one head x 128 query rows; KV tiles of 64 cells; 288 threads: warp 0 issues the MMAs, warps 1-8 stage the
//     operands, run the online softmax and keep the output rows in registers (thread = query row x half of the head dims);
//   * S = Q K^T: Q rows as the M = 128 operand (f16 K: Q rounded to f16 like the CPU; quantised K: Q quantised to q8_0 exactly like
//     quantize_row_q8_0, then d*q -- 18 significant bits -- split into an f16 hi + lo pair), K tile as the N = 64 operand staged
//     K-major ([cell][dim], f16 copied, q8_0 / q4_0 dequantised to d*q as hi + lo).  hi*hi + hi*lo + lo*hi products are exact, so
//     the quantised path reproduces the CPU's integer block dots to 2^-22;
//   * softmax in f32 straight from TMEM (tcgen05.ld: one S row per thread pair), mask added per element, running max / sum per
//     row, fully masked tiles skipped (one mask pass per tile decides liveness for the CTA);
//   * O_tile = P V: P as an f16 hi + lo pair (22 bits) in the K-major A layout, V staged MN-major ([cell][dim] rows are the
//     natural cache layout: no transpose), quantised V as hi + lo; the per-tile product is read back from TMEM and folded into the
//     register accumulators with the rescale factor exp(m_old - m_new).
// Roofline: tensor pipe (f16); algorithmic work 4 * H * n_q * n_kv_live * D flops.
#include "common.cuh"
#include "tc05.cuh"
#include "fattn_tc.h"

namespace {

enum { KV_F16 = 0, KV_Q8_0 = 1, KV_Q4_0 = 2 };
constexpr int HD = 128, TQ = 128, TC = 64;         // head size, query rows per CTA, cells per KV tile
constexpr int FT_THREADS = 9 * 32;
constexpr uint32_t LBO = 128;
constexpr uint32_t SBO_D = (HD / 8) * 128;         // Q and K tiles: the K dimension is the head dim (16 chunks of 16 bytes per row)
constexpr uint32_t SBO_C = (TC / 8) * 128;         // P tile (K dimension = cells) and the MN-major V tile (dim groups 1024 bytes apart)
constexpr uint32_t Q_BYTES = TQ * HD * 2, KV_BYTES = TC * HD * 2, P_BYTES = TQ * TC * 2;

struct FtShared {
    float xmax[2][TQ], xsum[2][TQ];
    uint64_t s_bar, o_bar;
    uint32_t tmem_base, pad;
};

__device__ __forceinline__ void ft_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();        // a hang on the GPU box costs a whole lease
    }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ bool ft_elect() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void split_h(float x, __half &hi, __half &lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}
// 8 floats -> one 16-byte chunk of f16 (hi) and, if LO, the chunk of the remainders
template <bool LO>
__device__ __forceinline__ void store_chunk(const float (&v)[8], uint8_t *hi_dst, uint8_t *lo_dst) {
    __half h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (LO) split_h(v[i], h[i], l[i]);
        else h[i] = __float2half_rn(v[i]);
    }
    *(uint4 *)hi_dst = *(const uint4 *)h;
    if (LO) *(uint4 *)lo_dst = *(const uint4 *)l;
}

// 32 consecutive head dims [32 qd, 32 qd + 32) of one cache row -> f32 (the values the CPU's dequantize_row_* / fp16 conversion gives)
template <int KT>
__device__ __forceinline__ void load_row32(const char *row, int qd, float (&v)[32]) {
    if (KT == KV_F16) {
        const uint4 *p = (const uint4 *)(row + qd * 64);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint4 w = p[j];
            const __half2 *h = (const __half2 *)&w;
#pragma unroll
            for (int e = 0; e < 4; e++) { const float2 f = __half22float2(h[e]); v[8 * j + 2 * e] = f.x; v[8 * j + 2 * e + 1] = f.y; }
        }
    } else if (KT == KV_Q8_0) {
        const unsigned short *p = (const unsigned short *)(row + qd * 34);         // block_q8_0: fp16 d, 32 int8 (2-byte aligned)
        const float d = __half2float(__ushort_as_half(p[0]));
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const unsigned short w = p[1 + i];
            v[2 * i] = d * (float)(int8_t)(w & 0xff);
            v[2 * i + 1] = d * (float)(int8_t)(w >> 8);
        }
    } else {
        const unsigned short *p = (const unsigned short *)(row + qd * 18);         // block_q4_0: fp16 d, 16 bytes: element i low nibble, i + 16 high nibble
        const float d = __half2float(__ushort_as_half(p[0]));
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const unsigned short w = p[1 + i];
            v[2 * i] = d * (float)((int)(w & 0xf) - 8);
            v[2 * i + 1] = d * (float)((int)((w >> 8) & 0xf) - 8);
            v[16 + 2 * i] = d * (float)((int)((w >> 4) & 0xf) - 8);
            v[16 + 2 * i + 1] = d * (float)((int)(w >> 12) - 8);
        }
    }
}

// ---- operand conversion without the f32 round trip -------------------------------------------------------------------------------
// The 32 head dims [32 qd, 32 qd + 32) of one raw cache row (shared memory, as stored) -> four 16-byte operand chunks of f16 `hi`
// and, for quantised caches, of the remainders `lo`.  Quantised: the value is d * q with d an f16 and q a small integer, so
//   hi = rn_f16(d * q) = HMUL2(d, q),   lo = rn_f16(d * q - hi) = HFMA2(d, q, -hi)   (the residual of a product is formed exactly)
// are bit for bit what split_h(d * (float)q) gives, at ~1 instruction per value instead of ~8.  Integers reach f16 through the
// 0x6400 magic number (PRMT a byte under the exponent of 1024, subtract 1024 + bias).  Blocks are 2-byte aligned: words are fetched
// aligned and funnel-shifted.
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *(const uint32_t *)&h; }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *(const __half2 *)&u; }
template <int KT>
__device__ __forceinline__ void convert_row32(const uint8_t *row, int qd, uint4 (&hi)[4], uint4 (&lo)[4]) {
    if (KT == KV_F16) {
        const uint4 *p = (const uint4 *)(row + qd * 64);
#pragma unroll
        for (int j = 0; j < 4; j++) hi[j] = p[j];
        return;
    }
    constexpr int BB = KT == KV_Q8_0 ? 34 : 18, NW = KT == KV_Q8_0 ? 9 : 5;
    const uint8_t *b = row + qd * BB;
    const uint32_t mis = (uint32_t)(uintptr_t)b & 2u;                    // the block's 2-byte phase inside its first aligned word
    const uint32_t *wp = (const uint32_t *)(b - mis);
    uint32_t w[NW + 1];
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = wp[i];
    w[NW] = 0;
    // q[i]: the quant words (bytes 2.. of the block); d: the block scale
    const uint32_t sh = mis ? 0u : 16u;                                  // mis == 2: d is the high half of w[0], the quants start at w[1]
    const uint32_t dbits = mis ? (w[0] >> 16) : (w[0] & 0xffffu);
    const __half2 d2 = u2h(dbits | (dbits << 16));
    constexpr int NQ = KT == KV_Q8_0 ? 8 : 4;
    uint32_t q[NQ];
#pragma unroll
    for (int i = 0; i < NQ; i++) q[i] = mis ? w[i + 1] : __funnelshift_r(w[i], w[i + 1], sh);
    auto emit = [&](uint32_t bytes4, __half2 bias, uint32_t &h0, uint32_t &h1, uint32_t &l0, uint32_t &l1) {
        // 4 unsigned bytes -> 4 integers (byte - bias) -> d * q as hi + lo
        const __half2 a = __hsub2(u2h(__byte_perm(bytes4, 0x64646464u, 0x4140)), bias), c = __hsub2(u2h(__byte_perm(bytes4, 0x64646464u, 0x4342)), bias);
        const __half2 ha = __hmul2(d2, a), hc = __hmul2(d2, c);
        h0 = h2u(ha); h1 = h2u(hc);
        l0 = h2u(__hfma2(d2, a, __hneg2(ha))); l1 = h2u(__hfma2(d2, c, __hneg2(hc)));
    };
    if (KT == KV_Q8_0) {
        const __half2 bias = __float2half2_rn(1024.0f + 128.0f);         // int8 + 128 = unsigned byte
#pragma unroll
        for (int j = 0; j < 4; j++) {
            emit(q[2 * j] ^ 0x80808080u, bias, hi[j].x, hi[j].y, lo[j].x, lo[j].y);
            emit(q[2 * j + 1] ^ 0x80808080u, bias, hi[j].z, hi[j].w, lo[j].z, lo[j].w);
        }
    } else {
        const __half2 bias = __float2half2_rn(1024.0f + 8.0f);           // block_q4_0: element i = low nibble of byte i, i + 16 = high nibble, minus 8
#pragma unroll
        for (int j = 0; j < 2; j++) {
            emit(q[2 * j] & 0x0f0f0f0fu, bias, hi[j].x, hi[j].y, lo[j].x, lo[j].y);
            emit(q[2 * j + 1] & 0x0f0f0f0fu, bias, hi[j].z, hi[j].w, lo[j].z, lo[j].w);
            emit((q[2 * j] >> 4) & 0x0f0f0f0fu, bias, hi[2 + j].x, hi[2 + j].y, lo[2 + j].x, lo[2 + j].y);
            emit((q[2 * j + 1] >> 4) & 0x0f0f0f0fu, bias, hi[2 + j].z, hi[2 + j].w, lo[2 + j].z, lo[2 + j].w);
        }
    }
}

// bytes of one head's row in the cache: 128 f16 / 4 q8_0 blocks / 4 q4_0 blocks
__host__ __device__ constexpr int raw_row_bytes(int kt) { return kt == KV_F16 ? HD * 2 : kt == KV_Q8_0 ? (HD / 32) * 34 : (HD / 32) * 18; }

// the 64 cache rows of KV tile t (one head) -> raw[cell][raw_row_bytes], by asynchronous copies from the 256 worker threads: the tile
// is on its way while the tensor core and the softmax work on the previous one.  f16 rows are 16-byte aligned, quantised rows 8-byte.
template <int KT>
__device__ __forceinline__ void fetch_tile(uint8_t *raw, const char *base, uint64_t nb1, int t, int wt) {
    constexpr int RB = raw_row_bytes(KT), PB = KT == KV_F16 ? 16 : 8, PC = RB / PB;      // pieces per row: 16 / 17 / 9
    const char *g = base + (uint64_t)t * TC * nb1;
    const uint32_t sb = smem_u32(raw);
    for (int i = wt; i < TC * PC; i += 256) {
        const int c = i / PC, pc = i - c * PC;
        if (KT == KV_F16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb + (uint32_t)(c * RB + pc * PB)), "l"(g + (uint64_t)c * nb1 + pc * PB) : "memory");
        else              asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sb + (uint32_t)(c * RB + pc * PB)), "l"(g + (uint64_t)c * nb1 + pc * PB) : "memory");
    }
}

template <int KT>
__global__ void __launch_bounds__(FT_THREADS, 1) b200_fattn_tc_kernel(const FaTcArgs p) {
    constexpr bool QUANT = KT != KV_F16;
    extern __shared__ __align__(1024) uint8_t fsm[];
    uint8_t *q_hi = fsm, *q_lo = q_hi + Q_BYTES;
    uint8_t *k_hi = q_lo + (QUANT ? Q_BYTES : 0), *k_lo = k_hi + KV_BYTES;
    uint8_t *v_hi = k_lo + (QUANT ? KV_BYTES : 0), *v_lo = v_hi + KV_BYTES;
    uint8_t *p_hi = v_lo + (QUANT ? KV_BYTES : 0), *p_lo = p_hi + P_BYTES;
    uint8_t *raw_k = p_lo + P_BYTES, *raw_v = raw_k + TC * raw_row_bytes(KT);          // the NEXT tile's cache rows, as stored (cp.async)
    FtShared *S = (FtShared *)(raw_v + TC * raw_row_bytes(KT));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = (int)gridDim.x - 1 - (int)blockIdx.x;       // late query blocks (most live tiles under a causal mask) first
    const int q0 = qb * TQ, h = blockIdx.y, hkv = h / p.gq;

    if (threadIdx.x == 0) { mbar_init(&S->s_bar, 1); mbar_init(&S->o_bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&S->tmem_base, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S->tmem_base;
    const uint32_t t_s = tmem, t_o = tmem + TC;                // S: 64 columns, O tile: 128 columns

    const int wt = (int)threadIdx.x - 32;                       // worker thread 0..255 (warps 1-8)
    // softmax / accumulate role: query row r (TMEM lane), half hf of the tile's cells and of the head dims
    const int lq = warp & 3, r = lq * 32 + lane, hf = warp >= 5 ? 1 : 0;
    const int q_idx = q0 + r;
    float out[HD / 2];
#pragma unroll
    for (int i = 0; i < HD / 2; i++) out[i] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f;

    if (warp > 0) {
        // ---- Q rows -> operand tile(s): thread = (row, half of the dims) ----
        const int rq = wt & (TQ - 1), hq = wt >> 7, qi = q0 + rq;
        const float *qp = (const float *)(p.q + (size_t)qi * p.q_nb1 + (size_t)h * p.q_nb2) + hq * 64;
#pragma unroll
        for (int b = 0; b < 2; b++) {                           // two 32-dim blocks (= q8_0 blocks of the row)
            float v[32];
            if (qi < p.n_q) {
#pragma unroll
                for (int j = 0; j < 8; j++) { const float4 f = *(const float4 *)(qp + 32 * b + 4 * j); v[4 * j] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w; }
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = 0.0f;
            }
            if (QUANT) {                                        // quantize_row_q8_0 (ggml-cpu-quants.c:808-845), then d * q (exact in f32)
                float amax = 0.0f;
#pragma unroll
                for (int j = 0; j < 32; j++) amax = fmaxf(amax, fabsf(v[j]));
                const float d = __fdiv_rn(amax, 127.0f);
                const float id = amax != 0.0f ? __fdiv_rn(127.0f, amax) : 0.0f;
                const float dq = __half2float(__float2half_rn(d));
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = dq * (float)__float2int_rn(__fmul_rn(v[j], id));
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float w[8] = {v[8 * c], v[8 * c + 1], v[8 * c + 2], v[8 * c + 3], v[8 * c + 4], v[8 * c + 5], v[8 * c + 6], v[8 * c + 7]};
                const uint32_t off = (uint32_t)(rq >> 3) * SBO_D + (uint32_t)(hq * 8 + b * 4 + c) * LBO + (uint32_t)(rq & 7) * 16;
                store_chunk<QUANT>(w, q_hi + off, q_lo + off);
            }
        }
    }

    const uint32_t idesc_s = (1u << 4) | ((uint32_t)(TC >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);                 // K-major A and B, N = 64
    const uint32_t idesc_o = (1u << 4) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);    // B MN-major, N = 128
    const uint64_t dd = make_desc(0, LBO, SBO_D), dc = make_desc(0, LBO, SBO_C);
    int live_tiles = 0;
    const int n_tiles = p.n_kv / TC;
    const char *kbase = p.k + (size_t)hkv * p.k_nb2, *vbase = p.v + (size_t)hkv * p.v_nb2;
    // mask values of (my row, my 32 cells of tile t); rows past n_q read as -inf
    uint4 mk[4];
    auto load_mask = [&](int t) {
        if (q_idx < p.n_q) {
            if (p.mask) {
                const uint4 *mp = (const uint4 *)(p.mask + (size_t)q_idx * p.m_nb1 + (size_t)(t * TC + hf * 32) * 2);
#pragma unroll
                for (int j = 0; j < 4; j++) mk[j] = mp[j];
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) mk[j] = make_uint4(0, 0, 0, 0);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) mk[j] = make_uint4(0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u);
        }
    };
    // software pipeline: tile t + 1's cache rows (cp.async -> raw_k / raw_v) and mask words (registers) are requested while tile t
    // is in the tensor core / softmax; ncu of the first version: long_scoreboard 6.1 stall cycles per issue, tensor pipe 8.5 % active
    if (warp > 0 && n_tiles > 0) {
        fetch_tile<KT>(raw_k, kbase, p.k_nb1, 0, wt);
        fetch_tile<KT>(raw_v, vbase, p.v_nb1, 0, wt);
        load_mask(0);
    }
    for (int t = 0; t < n_tiles; t++) {
        // ---- liveness of the whole tile (anything but -inf in any row) ----
        int live = 0;
        if (warp > 0) {
            if (q_idx < p.n_q) {
                if (p.mask) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t w[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
                        for (int e = 0; e < 4; e++) live |= ((w[e] & 0xffffu) != 0xfc00u) | ((w[e] >> 16) != 0xfc00u);      // anything but -inf
                    }
                } else live = 1;
            }
            cp_async_wait_all();                                // my pieces of tile t's rows have landed; the barrier below publishes them
        }
        if (!__syncthreads_or(live)) {                          // fully masked: nobody reads the raw rows; move on to the next tile
            if (warp > 0 && t + 1 < n_tiles) {
                fetch_tile<KT>(raw_k, kbase, p.k_nb1, t + 1, wt);
                fetch_tile<KT>(raw_v, vbase, p.v_nb1, t + 1, wt);
                load_mask(t + 1);
            }
            continue;
        }
        const uint32_t par = (uint32_t)(live_tiles & 1);
        live_tiles++;
        float alpha_t = 1.0f;

        if (warp > 0) {
            // ---- stage K (K-major) and V (MN-major): thread = (cell, quarter of the dims) ----
            const int c = wt & (TC - 1), qd = wt >> 6;
            uint4 ch[4], cl[4];
            convert_row32<KT>(raw_k + c * raw_row_bytes(KT), qd, ch, cl);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t off = (uint32_t)(c >> 3) * SBO_D + (uint32_t)(qd * 4 + j) * LBO + (uint32_t)(c & 7) * 16;
                *(uint4 *)(k_hi + off) = ch[j];
                if (QUANT) *(uint4 *)(k_lo + off) = cl[j];
            }
            convert_row32<KT>(raw_v + c * raw_row_bytes(KT), qd, ch, cl);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t off = (uint32_t)(qd * 4 + j) * SBO_C + (uint32_t)(c >> 3) * LBO + (uint32_t)(c & 7) * 16;
                *(uint4 *)(v_hi + off) = ch[j];
                if (QUANT) *(uint4 *)(v_lo + off) = cl[j];
            }
            fence_proxy_async();
        }
        __syncthreads();
        if (warp > 0 && t + 1 < n_tiles) {                      // every thread has converted its part: the raw buffers are free again
            fetch_tile<KT>(raw_k, kbase, p.k_nb1, t + 1, wt);
            fetch_tile<KT>(raw_v, vbase, p.v_nb1, t + 1, wt);
        }

        if (warp == 0) {
            // ---- S = Q K^T (+ Q K_lo^T + Q_lo K^T for quantised operands) ----
            tc_fence_after();
            if (ft_elect()) {
                const uint32_t qa = smem_u32(q_hi), ql = smem_u32(q_lo), ka = smem_u32(k_hi), kl = smem_u32(k_lo);
#pragma unroll
                for (int ks = 0; ks < HD / 16; ks++) mma_f16(t_s, dd + ((qa + ks * 2 * LBO) >> 4), dd + ((ka + ks * 2 * LBO) >> 4), idesc_s, ks ? 1u : 0u);
                if (QUANT) {
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ks++) mma_f16(t_s, dd + ((qa + ks * 2 * LBO) >> 4), dd + ((kl + ks * 2 * LBO) >> 4), idesc_s, 1u);
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ks++) mma_f16(t_s, dd + ((ql + ks * 2 * LBO) >> 4), dd + ((ka + ks * 2 * LBO) >> 4), idesc_s, 1u);
                }
                tc_commit(&S->s_bar);
            }
            __syncwarp();
        } else {
            // ---- online softmax of my 32 cells of row r ----
            ft_wait(&S->s_bar, par);
            tc_fence_after();
            uint32_t sv[32];
            {
                uint32_t a[16], b[16];
                const uint32_t ta = t_s + ((uint32_t)(lq * 32) << 16) + (uint32_t)(hf * 32);
                tmem_ld16(ta, a);
                tmem_ld16(ta + 16, b);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j++) { sv[j] = a[j]; sv[16 + j] = b[j]; }
            }
            float s[32], mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t w[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float2 mf = __half22float2(*(const __half2 *)&w[e]);
                    const int i = 8 * j + 2 * e;
                    s[i] = fmaf(__uint_as_float(sv[i]), p.scale, mf.x);
                    s[i + 1] = fmaf(__uint_as_float(sv[i + 1]), p.scale, mf.y);
                    mx = fmaxf(mx, fmaxf(s[i], s[i + 1]));
                }
            }
            if (t + 1 < n_tiles) load_mask(t + 1);              // mk is consumed: the next tile's words travel during the rest of this one
            S->xmax[hf][r] = mx;
            named_bar_sync(1, 256);
            const float m_new = fmaxf(m_run, fmaxf(mx, S->xmax[hf ^ 1][r]));
            float alpha = 1.0f, psum = 0.0f;
            if (m_new == -INFINITY) {
#pragma unroll
                for (int i = 0; i < 32; i++) s[i] = 0.0f;
            } else {
                alpha = expf(m_run - m_new);
#pragma unroll
                for (int i = 0; i < 32; i++) { s[i] = expf(s[i] - m_new); psum += s[i]; }
            }
            l_run = fmaf(l_run, alpha, psum);
            m_run = m_new;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float w[8] = {s[8 * j], s[8 * j + 1], s[8 * j + 2], s[8 * j + 3], s[8 * j + 4], s[8 * j + 5], s[8 * j + 6], s[8 * j + 7]};
                const uint32_t off = (uint32_t)(r >> 3) * SBO_C + (uint32_t)(hf * 4 + j) * LBO + (uint32_t)(r & 7) * 16;
                store_chunk<true>(w, p_hi + off, p_lo + off);
            }
            alpha_t = alpha;                                    // the rescale of the running output is folded into the accumulate below (one FFMA per element)
            fence_proxy_async();
            tc_fence_before();
        }
        __syncthreads();

        if (warp == 0) {
            // ---- O_tile = P_hi V + P_lo V (+ P_hi V_lo) ----
            tc_fence_after();
            if (ft_elect()) {
                const uint32_t pa = smem_u32(p_hi), pl = smem_u32(p_lo), va = smem_u32(v_hi), vl = smem_u32(v_lo);
#pragma unroll
                for (int ks = 0; ks < TC / 16; ks++) mma_f16(t_o, dc + ((pa + ks * 2 * LBO) >> 4), dc + ((va + ks * 2 * LBO) >> 4), idesc_o, ks ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < TC / 16; ks++) mma_f16(t_o, dc + ((pl + ks * 2 * LBO) >> 4), dc + ((va + ks * 2 * LBO) >> 4), idesc_o, 1u);
                if (QUANT) {
#pragma unroll
                    for (int ks = 0; ks < TC / 16; ks++) mma_f16(t_o, dc + ((pa + ks * 2 * LBO) >> 4), dc + ((vl + ks * 2 * LBO) >> 4), idesc_o, 1u);
                }
                tc_commit(&S->o_bar);
            }
            __syncwarp();
        } else {
            ft_wait(&S->o_bar, par);
            tc_fence_after();
            const uint32_t ta = t_o + ((uint32_t)(lq * 32) << 16) + (uint32_t)(hf * 64);
#pragma unroll
            for (int c0 = 0; c0 < HD / 2; c0 += 16) {
                uint32_t o[16];
                tmem_ld16(ta + c0, o);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j++) out[c0 + j] = fmaf(out[c0 + j], alpha_t, __uint_as_float(o[j]));
            }
            tc_fence_before();
        }
    }

    if (warp > 0) {
        S->xsum[hf][r] = l_run;
        named_bar_sync(1, 256);
        const float l = l_run + S->xsum[hf ^ 1][r];
        const float inv = __fdiv_rn(1.0f, l);
        if (q_idx < p.n_q) {
            float *dp = p.dst + ((size_t)q_idx * p.H + h) * HD + hf * 64;
#pragma unroll
            for (int j = 0; j < HD / 2; j += 4) *(float4 *)(dp + j) = make_float4(out[j] * inv, out[j + 1] * inv, out[j + 2] * inv, out[j + 3] * inv);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

template <int KT>
int launch_ft(b200_ctx *ctx, const FaTcArgs &p) {
    constexpr bool QUANT = KT != KV_F16;
    const size_t smem = (size_t)Q_BYTES * (QUANT ? 2 : 1) + (size_t)KV_BYTES * (QUANT ? 4 : 2) + 2 * P_BYTES + 2 * (size_t)TC * raw_row_bytes(KT) + sizeof(FtShared);
    auto kern = b200_fattn_tc_kernel<KT>;
    static bool attr_set[16] = {false};
    if (!attr_set[ctx->device & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        attr_set[ctx->device & 15] = true;
    }
    const dim3 grid((unsigned)((p.n_q + TQ - 1) / TQ), (unsigned)p.H);
    kern<<<grid, FT_THREADS, smem, ctx->stream>>>(p);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

}  // namespace

// kv_type: 0 f16, 1 q8_0, 2 q4_0 (K and V alike)
int fattn_tc_launch(b200_ctx *ctx, const FaTcArgs &p, int kv_type) {
    if (kv_type == KV_F16) return launch_ft<KV_F16>(ctx, p);
    if (kv_type == KV_Q8_0) return launch_ft<KV_Q8_0>(ctx, p);
    return launch_ft<KV_Q4_0>(ctx, p);
}
