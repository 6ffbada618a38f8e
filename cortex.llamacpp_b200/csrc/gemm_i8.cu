// gemm_i8.cu -- prefill GEMM for K-quant weights on the 5th-generation tensor cores (tcgen05.mma kind::i8, TMEM).
//
// Replaces mul_mat_q (ggml-cuda/mmq.cuh:2501-2657: mma.sync m16n8k32, q8_1 activations, float fix-up per 32-k block) with
// a B200 design whose integer stage is EXACTLY the CPU oracle's (ggml_vec_dot_q{4,5,6}_K_q8_K, ggml-cpu-quants.c):
//   * activations are quantised once per matmul to q8_K (quantize_row_q8_K_ref semantics, bit-exact) by a pre-pass that
//     writes the int8 quants directly as 128x128 tiles in the tensor core's canonical K-major shared-memory layout,
//     plus the per-256 scale d and the per-32 / per-16 quant sums the min / offset terms need;
//   * weights stay in their GGUF block layout in HBM.  A producer warp stages the raw blocks of a 128-row tile with bulk
//     async copies (TMA, cp.async.bulk + mbarrier); converter threads (one weight row each) expand a block to the integers
//     w' = sc_j * q (Q4_K/Q5_K: 6-bit sub-scale times 4/5-bit quant, <= 1953) or w' = scale_j * q6 (Q6_K), split them as
//     w' = 128*hi + lo into two int8 operand tiles in shared memory.  Per 256-element super-block the tensor core runs
//     2 x 8 tcgen05.mma (M=128 rows, N=128 tokens, K=32) into two int32 TMEM accumulators, so
//         P_b = 128 * sum(hi*a) + sum(lo*a) = sum_j sc_j * sum_l q_jl a_jl
//     is the CPU's per-block integer, bit for bit, with ONE accumulator drain per 256 k instead of one per 32;
//   * the epilogue threads (the same two warpgroups, alternating super-blocks) read the accumulators with tcgen05.ld,
//     add the min term M_b = sum_j m_j * (sum of a over sub-block j) (dp2a on the pre-computed sums) or the Q6_K offset
//     term, and accumulate  d_w*d_a*P_b - dmin_w*d_a*M_b  in f32 registers (128 per thread: one row x 128 tokens).
// Roofline: tensor pipe (int8).  Algorithmic work per launch: 2*N*K*M integer MACs-as-flops.
#include "common.cuh"
#include "gemm_i8.h"
#include "tc05.cuh"
#include "quant_warp.cuh"

namespace {

constexpr int TM = 128;            // weight rows per CTA tile  (MMA M)
constexpr int TN = 128;            // tokens per CTA tile       (MMA N)
constexpr int KH = 128;            // k per half-stage (half a super-block)
constexpr int NSLOT = 3;           // operand ring depth (half-stages)
constexpr int HALF_BYTES = TM * KH;                 // 16 KB per int8 operand tile
constexpr int GEMM_THREADS = 384;  // warp 0 weight producer, 1 MMA issuer, 2 TMEM owner, 3 activation producer, 4-7 / 8-11 warpgroups

// canonical K-major, no-swizzle operand layout (cute UMMA "INTERLEAVE": ((8,m),(16B,2)) : ((16B,SBO),(1,LBO))):
// 8 rows x 16 bytes form one 128-byte core matrix; core matrices adjacent in K are LBO apart, adjacent in M/N are SBO apart
constexpr uint32_t LBO = 128, SBO = (KH / 16) * 128;          // 128 B, 1024 B
__host__ __device__ __forceinline__ uint32_t canon_off(int r, int k) { return (uint32_t)(r >> 3) * SBO + (uint32_t)(k >> 4) * LBO + (uint32_t)(r & 7) * 16 + (uint32_t)(k & 15); }

struct GemmParams {
    const uint8_t *W; uint32_t rb; int type; int N, K, M;
    const uint8_t *Bq;             // [tok tile][K/128][16 KB]
    const float *Bd;               // [K/256][Mpad]
    const int16_t *Bs32;           // [K/256][Mpad][8]
    const int16_t *Bs16;           // [K/256][Mpad][16]
    int Mpad;
    float *dst; size_t dst_stride; // dst[tok * dst_stride + row]
    int raw_slots; uint32_t raw_row_bytes;     // per-row slot in the raw ring
    int swap_lbo_sbo;
    int sb_per_split;              // split-K: blockIdx.z covers super-blocks [z*sb_per_split, ...); partial results go to dst_partial
    float *dst_partial;            // [ksplit][M][N] (row stride N) when gridDim.z > 1
};

__device__ __forceinline__ int dp2a_lo_(uint32_t a, uint32_t b, int c) { return __dp2a_lo((int)a, (int)b, c); }
__device__ __forceinline__ int dp2a_hi_(uint32_t a, uint32_t b, int c) { return __dp2a_hi((int)a, (int)b, c); }

// ---- activation pre-pass: f32 [K, M] -> q8_K quants in canonical tiles + scales + sub-block sums ---------------------
// one warp per (token, super-block); lane owns 8 consecutive elements (quantize_row_q8_K_ref, ggml-quants.c:2479-2513)
__global__ void __launch_bounds__(128) b200_gemm_quantize_kernel(const float *__restrict__ x, size_t x_col_stride, int K, int M, int Mpad,
                                                                 uint8_t *__restrict__ Bq, float *__restrict__ Bd, int16_t *__restrict__ Bs32,
                                                                 int16_t *__restrict__ Bs16) {
    const int lane = threadIdx.x & 31;
    const int nsb = K / 256;
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= (int64_t)nsb * Mpad) return;
    const int tok = (int)(gw / nsb), b = (int)(gw % nsb);
    float v[8];
    if (tok < M) {
        const float *xp = (const float *)((const char *)x + (size_t)tok * x_col_stride) + b * 256 + lane * 8;
        const float4 a = *(const float4 *)xp, c = *(const float4 *)(xp + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = 0.0f;
    }
    uint2 qp; float d; int pair;
    warp_quant_q8k(v, lane, qp, d, pair);              // pair: valid in even lanes = sum of the 16-group lane/2
    const int k = lane * 8;                            // within the super-block
    const int g = 2 * b + (k >> 7), kk = k & 127;
    const int tile = tok >> 7, n = tok & 127;
    uint8_t *tp = Bq + ((size_t)tile * (K / KH) + g) * HALF_BYTES;
    *(uint2 *)(tp + canon_off(n, kk)) = qp;
    const int p2 = pair + __shfl_down_sync(0xffffffffu, pair, 2);      // lanes 0,4,8,..: sum of 32 elements
    const size_t row = (size_t)b * Mpad + tok;
    if ((lane & 1) == 0) Bs16[row * 16 + (lane >> 1)] = (int16_t)pair;
    if ((lane & 3) == 0) Bs32[row * 8 + (lane >> 2)] = (int16_t)p2;
    if (lane == 0) Bd[row] = d;
}

// ---- weight expansion: one thread = one weight row, one half super-block (128 weights) -> hi / lo int8 operand rows ----
struct RowMeta {                     // per (row, super-block) scale data kept in registers from conversion to epilogue
    float d, dmin;
    uint32_t m0123, m4567;           // Q4_K/Q5_K: 6-bit mins of the 8 sub-blocks; Q6_K: unused
    uint32_t sc[4];                  // Q6_K: the 16 int8 sub-block scales
};

__device__ __forceinline__ void get_scale_min(const uint8_t *scales, int j, int &sc, int &mn) {   // get_scale_min_k4, ggml-quants.c:631
    if (j < 4) { sc = scales[j] & 63; mn = scales[j + 4] & 63; }
    else { sc = (scales[j + 4] & 0x0f) | ((scales[j - 4] >> 6) << 4); mn = (scales[j + 4] >> 4) | ((scales[j] >> 6) << 4); }
}

// four products sc*q (q in the 4 bytes of `w`, <= 31; sc <= 63) -> hi bytes and lo bytes
__device__ __forceinline__ void split4(uint32_t w, uint32_t sc, uint32_t &hi, uint32_t &lo) {
    const uint32_t e = (w & 0x00ff00ffu) * sc, o = ((w >> 8) & 0x00ff00ffu) * sc;       // two 16-bit lanes each, no carry (<= 1953)
    hi = ((e >> 7) & 0x00ff00ffu) | (((o >> 7) & 0x00ff00ffu) << 8);
    lo = (e & 0x007f007fu) | ((o & 0x007f007fu) << 8);
}

template <int TYPE>
__device__ __forceinline__ void read_meta(const uint8_t *blk, RowMeta &m) {
    if (TYPE == B200_TYPE_Q6_K) {
        m.d = __half2float(*(const __half *)(blk + 208));
        m.dmin = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) m.sc[i] = (uint32_t)blk[192 + 4 * i] | (uint32_t)blk[193 + 4 * i] << 8 | (uint32_t)blk[194 + 4 * i] << 16 | (uint32_t)blk[195 + 4 * i] << 24;
        m.m0123 = m.m4567 = 0;
    } else {
        m.d = __half2float(*(const __half *)blk);
        m.dmin = __half2float(*(const __half *)(blk + 2));
        uint32_t mm[2] = {0, 0};
#pragma unroll
        for (int j = 0; j < 8; j++) { int sc, mn; get_scale_min(blk + 4, j, sc, mn); mm[j >> 2] |= (uint32_t)mn << (8 * (j & 3)); }
        m.m0123 = mm[0]; m.m4567 = mm[1];
    }
}

// writes the 128 bytes of operand row r for half h into the hi and lo tiles (canonical layout: 8 chunks of 16 bytes)
template <int TYPE>
__device__ __forceinline__ void convert_half(const uint8_t *blk, int h, int r, uint8_t *a_hi, uint8_t *a_lo) {
    if (TYPE == B200_TYPE_Q4_K || TYPE == B200_TYPE_Q5_K) {
        const uint8_t *qs = blk + (TYPE == B200_TYPE_Q5_K ? 48 : 16);
        const uint8_t *qh = blk + 16;
#pragma unroll
        for (int gi = 0; gi < 2; gi++) {                      // two 32-byte qs groups per half: sub-blocks (4h+2gi, 4h+2gi+1)
            const int g = 2 * h + gi;
            int sc0, m0, sc1, m1;
            get_scale_min(blk + 4, 2 * g, sc0, m0);
            get_scale_min(blk + 4, 2 * g + 1, sc1, m1);
            uint32_t q[8];
#pragma unroll
            for (int i = 0; i < 8; i++) q[i] = *(const uint32_t *)(qs + 32 * g + 4 * i);
            uint32_t hb[8];
            if (TYPE == B200_TYPE_Q5_K) {
#pragma unroll
                for (int i = 0; i < 8; i++) hb[i] = *(const uint32_t *)(qh + 4 * i);
            }
#pragma unroll
            for (int sub = 0; sub < 2; sub++) {               // low nibbles = sub-block 2g, high nibbles = 2g+1
                const uint32_t sc = sub ? (uint32_t)sc1 : (uint32_t)sc0;
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    uint32_t w = sub ? (q[i] >> 4) & 0x0f0f0f0fu : q[i] & 0x0f0f0f0fu;
                    if (TYPE == B200_TYPE_Q5_K) w |= ((hb[i] >> (2 * g + sub)) & 0x01010101u) << 4;
                    split4(w, sc, hi[i], lo[i]);
                }
                const int k0 = (2 * gi + sub) * 32;          // k offset of this sub-block inside the half
                *(uint4 *)(a_hi + canon_off(r, k0))      = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *(uint4 *)(a_hi + canon_off(r, k0 + 16)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                *(uint4 *)(a_lo + canon_off(r, k0))      = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                *(uint4 *)(a_lo + canon_off(r, k0 + 16)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
        }
    } else {
        // Q6_K: half h = elements [128h, 128h+128): ql[64h..64h+63], qh[32h..32h+31], scales[8h..8h+7] (ggml-quants.c:1690-1719).
        // Signed quants v = q - 32 in [-32, 31] times the int8 scale: a = sc * v in [-4064, 4096] = 128 * hi + lo with lo in [0, 127] and
        // hi = a >> 7 in [-32, 32]: the CPU's integer sum_j scale_j * sum (q - 32) a comes out of the tensor core directly, no offset term.
        // Word-wise: 4 elements per 32-bit word of ql / qh (blocks are only 2-byte aligned: 16-bit loads).
        const uint8_t *ql = blk + 64 * h, *qh = blk + 128 + 32 * h;
        const int8_t *scp = (const int8_t *)(blk + 192 + 8 * h);
        auto ld32 = [](const uint8_t *p) { return (uint32_t)*(const unsigned short *)p | ((uint32_t)*(const unsigned short *)(p + 2) << 16); };
        uint32_t qhw[8];
#pragma unroll
        for (int i = 0; i < 8; i++) qhw[i] = ld32(qh + 4 * i);
#pragma unroll
        for (int t = 0; t < 4; t++) {                         // quadrant t: elements 32t .. 32t+31 of the half
#pragma unroll
            for (int c16 = 0; c16 < 2; c16++) {              // 16-byte chunk of the quadrant = one 16-element scale group
                const int sc = scp[2 * t + c16];
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int wd = 0; wd < 4; wd++) {
                    const int l = 16 * c16 + 4 * wd;          // byte index 0..31 inside the quadrant
                    const uint32_t lw = ld32(ql + l + 32 * (t & 1));
                    const uint32_t q4 = ((t < 2 ? lw : lw >> 4) & 0x0f0f0f0fu) | (((qhw[l >> 2] >> (2 * t)) & 0x03030303u) << 4);    // 4 x (0..63)
                    // two 16-bit lanes per multiply: (q - 32) * sc as signed 16-bit lanes of a 32-bit product is not carry-free, so go through
                    // the unsigned product and subtract 32 * sc per lane: a = q * sc - 32 * sc, each lane handled as a signed 16-bit value
                    const int e0 = (int)(q4 & 0xffu) * sc - 32 * sc, e1 = (int)((q4 >> 8) & 0xffu) * sc - 32 * sc;
                    const int e2 = (int)((q4 >> 16) & 0xffu) * sc - 32 * sc, e3 = (int)(q4 >> 24) * sc - 32 * sc;
                    hi[wd] = (uint32_t)((e0 >> 7) & 0xff) | (uint32_t)((e1 >> 7) & 0xff) << 8 | (uint32_t)((e2 >> 7) & 0xff) << 16 | (uint32_t)((e3 >> 7) & 0xff) << 24;
                    lo[wd] = (uint32_t)(e0 & 127) | (uint32_t)(e1 & 127) << 8 | (uint32_t)(e2 & 127) << 16 | (uint32_t)(e3 & 127) << 24;
                }
                const int k0 = 32 * t + 16 * c16;
                *(uint4 *)(a_hi + canon_off(r, k0)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *(uint4 *)(a_lo + canon_off(r, k0)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
}

// ---- the GEMM kernel --------------------------------------------------------------------------------------------------
struct SmemLayout {
    uint64_t raw_full[4], raw_empty[4], a_full[NSLOT], ab_empty[NSLOT], b_full[NSLOT], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    uint32_t pad[31];
};

// PARTIAL: the token tile may hold fewer than TN real tokens (decode-sized ubatches): drain only the accumulator columns in use
template <int TYPE, bool PARTIAL>
__global__ void __launch_bounds__(GEMM_THREADS, 1) b200_gemm_i8_kernel(const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t gsm[];
    uint8_t *a_hi = gsm;                                   // [NSLOT][16 KB]
    uint8_t *a_lo = a_hi + NSLOT * HALF_BYTES;
    uint8_t *bq   = a_lo + NSLOT * HALF_BYTES;
    uint8_t *raw  = bq + NSLOT * HALF_BYTES;               // [raw_slots][TM][raw_row_bytes]
    SmemLayout *S = (SmemLayout *)(raw + (size_t)p.raw_slots * TM * p.raw_row_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * TM, tok0 = blockIdx.y * TN;
    constexpr int BLK = TYPE == B200_TYPE_Q4_K ? 144 : TYPE == B200_TYPE_Q5_K ? 176 : 210;
    // split-K (small token counts leave most SMs idle otherwise): this CTA sees a K-slice as a GEMM of its own over shifted views
    const int sb0 = (int)blockIdx.z * p.sb_per_split;
    const int nsb = min(p.sb_per_split, p.K / 256 - sb0);
    const uint8_t *const Wv = p.W + (size_t)sb0 * BLK;
    const uint8_t *const Bqv = p.Bq + (size_t)(2 * sb0) * HALF_BYTES;
    const float *const Bdv = p.Bd + (size_t)sb0 * p.Mpad;
    const int16_t *const Bs32v = p.Bs32 + (size_t)sb0 * p.Mpad * 8;
    const int16_t *const Bs16v = p.Bs16 + (size_t)sb0 * p.Mpad * 16;
    float *const dstv = gridDim.z > 1 ? p.dst_partial + (size_t)blockIdx.z * p.M * p.N : p.dst;
    const size_t dst_stride = gridDim.z > 1 ? (size_t)p.N : p.dst_stride;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; i++) { mbar_init(&S->raw_full[i], 1); mbar_init(&S->raw_empty[i], 4); }
        for (int i = 0; i < NSLOT; i++) { mbar_init(&S->a_full[i], 4); mbar_init(&S->ab_empty[i], 1); mbar_init(&S->b_full[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&S->acc_full[i], 1); mbar_init(&S->acc_empty[i], 4); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(&S->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S->tmem_base;

    if (warp == 0) {
        // ------------------------------------------------------------ producer: raw weight blocks (TMA bulk copies)
        // lane l stages rows l, l+32, l+64, l+96 of the tile: one block each per super-block
        for (int b = 0; b < nsb; b++) {
            const int rs = b % p.raw_slots;
            if (b >= p.raw_slots) mbar_wait(&S->raw_empty[rs], ((b / p.raw_slots) - 1) & 1);
            uint32_t mybytes = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int r = lane + 32 * i;
                const uint8_t *src = Wv + (size_t)(row0 + r) * p.rb + (size_t)b * BLK;
                const uint32_t extra = (uint32_t)((uintptr_t)src & 15);
                mybytes += (extra + BLK + 15u) & ~15u;
            }
            uint32_t total = mybytes;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            if (lane == 0) mbar_arrive_expect_tx(&S->raw_full[rs], total);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int r = lane + 32 * i;
                const uint8_t *src = Wv + (size_t)(row0 + r) * p.rb + (size_t)b * BLK;
                const uint32_t extra = (uint32_t)((uintptr_t)src & 15);
                const uint32_t bytes = (extra + BLK + 15u) & ~15u;
                bulk_g2s(raw + ((size_t)rs * TM + r) * p.raw_row_bytes, src - extra, bytes, &S->raw_full[rs]);
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------ producer: activation tiles (one 16 KB bulk copy per half-stage)
        if (lane == 0) {
            const uint8_t *src = Bqv + (size_t)blockIdx.y * (p.K / KH) * HALF_BYTES;
            for (int g = 0; g < 2 * nsb; g++) {
                const int slot = g % NSLOT;
                if (g >= NSLOT) mbar_wait(&S->ab_empty[slot], ((g / NSLOT) - 1) & 1);
                mbar_arrive_expect_tx(&S->b_full[slot], HALF_BYTES);
                bulk_g2s(bq + (size_t)slot * HALF_BYTES, src + (size_t)g * HALF_BYTES, HALF_BYTES, &S->b_full[slot]);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            const uint32_t lbo = p.swap_lbo_sbo ? SBO : LBO, sbo = p.swap_lbo_sbo ? LBO : SBO;
            for (int g = 0; g < 2 * nsb; g++) {
                const int b = g >> 1, h = g & 1, pr = b & 1, slot = g % NSLOT;
                if (h == 0 && b >= 2) mbar_wait(&S->acc_empty[pr], ((b >> 1) - 1) & 1);
                mbar_wait(&S->a_full[slot], (g / NSLOT) & 1);
                mbar_wait(&S->b_full[slot], (g / NSLOT) & 1);
                tc_fence_after();
                const uint32_t ah = smem_u32(a_hi + (size_t)slot * HALF_BYTES), al = smem_u32(a_lo + (size_t)slot * HALF_BYTES);
                const uint32_t bb = smem_u32(bq + (size_t)slot * HALF_BYTES);
                const uint32_t d_hi = tmem + (uint32_t)pr * 256, d_lo = d_hi + 128;
#pragma unroll
                for (int ks = 0; ks < KH / 32; ks++) {
                    const uint32_t acc = (h | ks) ? 1u : 0u;
                    const uint64_t bd = make_desc(bb + ks * 2 * LBO, lbo, sbo);
                    mma_i8(d_hi, make_desc(ah + ks * 2 * LBO, lbo, sbo), bd, idesc, acc);
                    mma_i8(d_lo, make_desc(al + ks * 2 * LBO, lbo, sbo), bd, idesc, acc);
                }
                tc_commit(&S->ab_empty[slot]);             // operand slot free once these MMAs have read it
                if (h == 1) tc_commit(&S->acc_full[pr]);   // both halves accumulated: the super-block's integers are ready
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ warpgroups: expand weights, then drain accumulators
        const int wg = (warp - 4) >> 2;                    // 0 / 1: owns super-blocks b = wg (mod 2) and accumulator pair wg
        const int r = (warp & 3) * 32 + lane;              // weight row in the tile == TMEM lane
        float out[TN];
#pragma unroll
        for (int c = 0; c < TN; c++) out[c] = 0.0f;
        RowMeta prev, cur;
        int pb = -1;
        const int ncols_valid = min(TN, p.M - tok0);
        const int last_c0 = ((ncols_valid + 31) / 32 - 1) * 32;     // last 32-column chunk of the accumulators that holds real tokens

        auto epilogue = [&](int b, const RowMeta &m) {
            const int pr = b & 1;
            mbar_wait(&S->acc_full[pr], (b >> 1) & 1);
            tc_fence_after();
            const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32) << 16);
            const size_t rowb = (size_t)b * p.Mpad + tok0;
            const float dd = m.d, ndm = -m.dmin;
#pragma unroll
            for (int c0 = 0; c0 < TN; c0 += 32) {
                if (PARTIAL && c0 > last_c0) continue;     // token columns beyond M (a 32-slot decode ubatch fills a quarter of the tile)
                uint32_t hi[32], lo[32];
                tmem_ld32(tmem + lane_addr + (uint32_t)pr * 256 + c0, hi);
                tmem_ld32(tmem + lane_addr + (uint32_t)pr * 256 + 128 + c0, lo);
                tmem_ld_wait();
                if (PARTIAL ? c0 == last_c0 : c0 == TN - 32) { // accumulators are in registers: the tensor core may overwrite them
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S->acc_empty[pr]);
                }
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int c = c0 + j;
                    int P = ((int)hi[j] << 7) + (int)lo[j];
                    const float da = __ldg(Bdv + rowb + c);
                    if (TYPE == B200_TYPE_Q6_K) {
                        out[c] = fmaf(dd * da, (float)P, out[c]);          // signed quants: no offset term (convert_half)
                    } else {
                        const uint4 s = __ldg((const uint4 *)(Bs32v + (rowb + c) * 8));
                        int Mv = 0;
                        Mv = dp2a_lo_(s.x, m.m0123, Mv); Mv = dp2a_hi_(s.y, m.m0123, Mv);
                        Mv = dp2a_lo_(s.z, m.m4567, Mv); Mv = dp2a_hi_(s.w, m.m4567, Mv);
                        out[c] = fmaf(da, fmaf(dd, (float)P, ndm * (float)Mv), out[c]);
                    }
                }
            }
        };

        for (int b = wg;; b += 2) {
            const bool more = b < nsb;
            if (more) {
                const int rs = b % p.raw_slots;
                mbar_wait(&S->raw_full[rs], (b / p.raw_slots) & 1);
                const uint8_t *slotp = raw + ((size_t)rs * TM + r) * p.raw_row_bytes;
                const uint8_t *blk = slotp + (uint32_t)((uintptr_t)(Wv + (size_t)(row0 + r) * p.rb + (size_t)b * BLK) & 15);
                read_meta<TYPE>(blk, cur);
#pragma unroll 1
                for (int h = 0; h < 2; h++) {
                    const int g = 2 * b + h, slot = g % NSLOT;
                    if (g >= NSLOT) mbar_wait(&S->ab_empty[slot], ((g / NSLOT) - 1) & 1);
                    convert_half<TYPE>(blk, h, r, a_hi + (size_t)slot * HALF_BYTES, a_lo + (size_t)slot * HALF_BYTES);
                    fence_proxy_async();                   // generic-proxy stores -> visible to the tensor core's async proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S->a_full[slot]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&S->raw_empty[rs]);
            }
            if (pb >= 0) epilogue(pb, prev);               // drain the PREVIOUS super-block of this warpgroup while the tensor core works
            if (!more) break;
            prev = cur; pb = b;
        }

        // ---- combine the two warpgroups' partial sums (fixed order: wg0 + wg1) and store ----
        __syncwarp();
        // reuse the operand tiles as exchange space once every MMA has retired (all accumulators were drained above)
        asm volatile("bar.sync 2, 256;" ::: "memory");
        float *xch = (float *)gsm;                         // [TN][TM] floats = 64 KB
        if (wg == 1) {
#pragma unroll
            for (int c = 0; c < TN; c++) xch[c * TM + r] = out[c];
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (wg == 0) {
#pragma unroll
            for (int c = 0; c < TN; c++) {
                const int tok = tok0 + c;
                if (tok < p.M) dstv[(size_t)tok * dst_stride + row0 + r] = out[c] + xch[c * TM + r];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
}


// dst[tok][n] = sum over the K-slices in slice order (deterministic)
__global__ void b200_gemm_splitk_reduce_kernel(const float *__restrict__ part, int ksplit, int N, int M, float *__restrict__ dst, size_t dst_stride) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)M * N) return;
    const int tok = (int)(i / N), n = (int)(i % N);
    float a = part[i];
    for (int z = 1; z < ksplit; z++) a += part[(size_t)z * M * N + i];
    dst[(size_t)tok * dst_stride + n] = a;
}

template <int TYPE, bool PARTIAL>
int launch_gemm_t(b200_ctx *ctx, const GemmParams &p, dim3 grid, size_t smem) {
    auto kern = b200_gemm_i8_kernel<TYPE, PARTIAL>;
    static bool attr_set[16] = {false};
    if (!attr_set[ctx->device & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        attr_set[ctx->device & 15] = true;
    }
    kern<<<grid, GEMM_THREADS, smem, ctx->stream>>>(p);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

template <int TYPE>
int launch_gemm(b200_ctx *ctx, const GemmParams &p, dim3 grid, size_t smem) {
    return p.M % TN ? launch_gemm_t<TYPE, true>(ctx, p, grid, smem) : launch_gemm_t<TYPE, false>(ctx, p, grid, smem);
}

}  // namespace

int g_gemm_desc_swap = -1;     // debug: exchange the LBO/SBO fields of the shared-memory descriptors

bool gemm_i8_supported(int type, int64_t N, int64_t K, int64_t M) {
    if (type != B200_TYPE_Q4_K && type != B200_TYPE_Q5_K && type != B200_TYPE_Q6_K) return false;
    if (getenv("GGML_B200_NO_GEMM")) return false;
    return N % TM == 0 && K % 256 == 0 && M > 8;
}

// dst[tok * dst_stride + n] = sum_k W[n, k] * x[k, tok]; x f32 with column stride x_stride bytes
int gemm_i8_run(b200_ctx *ctx, int type, const uint8_t *W, size_t rb, int64_t N, int64_t K, const float *x, size_t x_stride, int64_t M,
                float *dst, size_t dst_stride) {
    if (!gemm_i8_supported(type, N, K, M)) { b200_set_error("gemm_i8: unsupported shape"); return B200_ERR_UNSUPPORTED; }
    const int ntile = (int)((M + TN - 1) / TN);
    const int Mpad = ntile * TN;
    const int nsb = (int)(K / 256);
    // scratch: canonical int8 tiles | d | sums per 32 | sums per 16
    const size_t sz_q = (size_t)ntile * (K / KH) * HALF_BYTES;
    const size_t sz_d = ((size_t)nsb * Mpad * 4 + 255) & ~(size_t)255;
    const size_t sz_s32 = ((size_t)nsb * Mpad * 8 * 2 + 255) & ~(size_t)255;
    const size_t sz_s16 = ((size_t)nsb * Mpad * 16 * 2 + 255) & ~(size_t)255;
    uint8_t *scr = (uint8_t *)ctx->get_scratch(SCRATCH_ACT, sz_q + sz_d + sz_s32 + sz_s16);
    if (!scr) return B200_ERR_ALLOC;
    GemmParams p = {};
    p.W = W; p.rb = (uint32_t)rb; p.type = type; p.N = (int)N; p.K = (int)K; p.M = (int)M;
    p.Bq = scr; p.Bd = (const float *)(scr + sz_q); p.Bs32 = (const int16_t *)(scr + sz_q + sz_d); p.Bs16 = (const int16_t *)(scr + sz_q + sz_d + sz_s32);
    p.Mpad = Mpad; p.dst = dst; p.dst_stride = dst_stride;
    if (g_gemm_desc_swap < 0) g_gemm_desc_swap = getenv("GGML_B200_GEMM_SWAP") ? atoi(getenv("GGML_B200_GEMM_SWAP")) : 0;
    p.swap_lbo_sbo = g_gemm_desc_swap;
    {
        const int64_t warps = (int64_t)nsb * Mpad;
        b200_gemm_quantize_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, ctx->stream>>>(x, x_stride, (int)K, (int)M, Mpad, scr, (float *)p.Bd,
                                                                                      (int16_t *)p.Bs32, (int16_t *)p.Bs16);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    const int blk = type == B200_TYPE_Q4_K ? 144 : type == B200_TYPE_Q5_K ? 176 : 210;
    p.raw_row_bytes = (uint32_t)((blk + 15 + 15) & ~15);            // worst-case misalignment + round-up
    if (type != B200_TYPE_Q6_K) p.raw_row_bytes = (uint32_t)blk;   // 16-byte multiples, 16-byte aligned rows
    const size_t fixed = (size_t)3 * NSLOT * HALF_BYTES + sizeof(SmemLayout) + 1024;
    int rs = (int)((ctx->smem_optin - fixed) / ((size_t)TM * p.raw_row_bytes));
    if (rs > 4) rs = 4;
    if (rs < 1) { b200_set_error("gemm_i8: shared memory"); return B200_ERR_FAILED; }
    p.raw_slots = rs;
    const size_t smem = (size_t)3 * NSLOT * HALF_BYTES + (size_t)rs * TM * p.raw_row_bytes + sizeof(SmemLayout);
    // split-K when the (row tile x token tile) grid leaves most of the machine idle; slices of >= 4 super-blocks
    static const int max_split = getenv("GGML_B200_GEMM_KSPLIT") ? atoi(getenv("GGML_B200_GEMM_KSPLIT")) : 8;
    const int64_t ctas = (N / TM) * ntile;
    int ksplit = 1;
    while (ksplit < max_split && ctas * ksplit * 2 <= ctx->sm_count && nsb / (ksplit * 2) >= 4) ksplit *= 2;
    p.sb_per_split = (nsb + ksplit - 1) / ksplit;
    ksplit = (nsb + p.sb_per_split - 1) / p.sb_per_split;
    if (ksplit > 1) {
        p.dst_partial = (float *)ctx->get_scratch(SCRATCH_MISC, (size_t)ksplit * M * N * 4);
        if (!p.dst_partial) return B200_ERR_ALLOC;
    }
    const dim3 grid((unsigned)(N / TM), (unsigned)ntile, (unsigned)ksplit);
    int rc;
    switch (type) {
        case B200_TYPE_Q4_K: rc = launch_gemm<B200_TYPE_Q4_K>(ctx, p, grid, smem); break;
        case B200_TYPE_Q5_K: rc = launch_gemm<B200_TYPE_Q5_K>(ctx, p, grid, smem); break;
        default:             rc = launch_gemm<B200_TYPE_Q6_K>(ctx, p, grid, smem); break;
    }
    if (rc || ksplit == 1) return rc;
    const int64_t total = M * N;
    b200_gemm_splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(p.dst_partial, ksplit, (int)N, (int)M, dst, dst_stride);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}
