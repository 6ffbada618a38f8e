// gemv_mma.cu -- small-batch matmul (5..32 token columns): the continuous-batching decode step (n_parallel slots, one token each).
//
// Replaces, for these batch sizes, mul_mat_vec_q with ncols_y > 1 (mmvq.cu:130-204, batch <= 8) and the small mmq_x tiles of
// mul_mat_q (mmq.cuh:2501-2657).  Round 1 sent 32 columns through the tcgen05 prefill GEMM, whose per-super-block drain costs the
// same for 32 tokens as for 128: 22.6 ms per bs32 step, 0.04 of the HBM roofline (profiles/r1_batched.md).  At this size the matmul
// is a weight-streaming problem like the batch-1 GEMV, only with 32x the integer work per weight byte, so it is built like the
// GEMV (persistent CTAs, a producer warp streaming GGUF rows through a shared-memory ring with cp.async.bulk + mbarrier) with
// the block dots on the warp-level tensor core path (mma.sync m16n8k32 s8: 4096 MACs per instruction; a 16 x 32 output tile
// does not fill a tcgen05 M=128 tile):
//   * unit of work = 16 weight rows x one K-range of 8 super-blocks (2048 k) = one ring stage (16 row pieces of 1-2 KB);
//   * CTA (g, r) keeps the quantised activations of K-range r resident in shared memory (<= 32 columns x 2048 int8, laid out
//     for ldmatrix) and walks the row tiles of group g; several K-ranges are combined by a fixed-order reduction (split-K);
//   * two warps share a stage and take alternate super-blocks of it: a warp expands the weight nibbles of a block to int8 A
//     fragments in registers ONCE (GGUF layout read in place) and multiplies them into all 8..32 token columns; per 32-wide
//     sub-block the int32 products are scaled by the 6-bit sub-scale (exactly the CPU's  sum_j sc_j sum_l q a), mins go through
//     dp2a on per-32 activation sums, q6_K / q4_0 / q8_0 use signed quants so they have no offset term; the two warps' float
//     partials are added in a fixed order through shared memory;
//   * the producer warp also brings the activation quants in with bulk copies (after griddepcontrol.wait; the weight ring is
//     filled before it, so the previous kernel's tail overlaps the first weight stages);
//   * activations are quantised exactly like the CPU (q8_K / q8_0, quant.cu) -> every per-block integer equals the oracle's.
// Roofline: HBM (weights once per step); algorithmic bytes = N*K*bpw + ncols*(K*1.14 + 4N).
#include "common.cuh"

int launch_quantize_act(b200_ctx *ctx, int q8k, const float *x, size_t x_col_stride_bytes, int64_t K, int64_t ncols, uint8_t *scratch);

namespace {

constexpr int MM_ROWS = 16;                 // rows per tile (mma M)
constexpr int MM_KB = 8;                    // 256-element blocks per K-range
constexpr int MM_KR = MM_KB * 256;          // 2048 k
constexpr int MM_ASTRIDE = MM_KR + 16;      // bytes per activation column in shared memory: (stride / 16) odd -> ldmatrix conflict-free
constexpr int MM_COLS = 32;
constexpr int MM_MAX_STAGES = 7;              // named barriers 2..8 and 9..15 belong to the warp pairs
constexpr int MM_THREADS = (2 * MM_MAX_STAGES + 1) * 32;

struct MmParams {
    const uint8_t *W; uint32_t rb; int type, N, K, ncols;
    const uint8_t *act; ActLayout L;
    float *dst; size_t dst_stride;
    const float *residual;                  // optional, laid out like dst (added once: in the epilogue or by the split-K reduction)
    float *part;                            // [nranges][MM_COLS][N] when nranges > 1
    int nranges, ngroups, ntiles, nb;       // nb: 256-blocks per row
    int nstages; uint32_t stage_bytes, rstride, bbytes;      // rstride: bytes between rows inside a stage; bbytes: bytes per 256 weights
    uint32_t off_aq, off_ad, off_as, off_comb, off_ring;
    int use_pdl, stream_once, w_const;
    // grouped (MUL_MAT_ID) mode: blockIdx.y = chunk of <= 32 (token, slot) pairs routed to one expert
    const int32_t *g_off;                   // [E + 1] first pair of every expert in g_pairs (device, written by the grouping kernel)
    const int32_t *g_pairs;                 // pair ids (token * n_used + slot) sorted by expert
    int g_E, g_n_used, g_b_ne1;             // g_E == 0: dense mode
    size_t g_expert_stride;                 // bytes between expert matrices
    size_t g_d_nb1, g_d_nb2;                // dst element strides of slot and token
};

// grouped mode: which (expert, chunk) does blockIdx.y stand for?  cnt = pairs in this chunk (0: nothing to do)
__device__ __forceinline__ void group_lookup(const MmParams &p, int y, int &e, int &first, int &cnt) {
    cnt = 0; first = 0;
    for (e = 0; e < p.g_E; e++) {
        const int o0 = p.g_off[e], n = p.g_off[e + 1] - o0, nch = (n + MM_COLS - 1) / MM_COLS;
        if (y < nch) { first = o0 + y * MM_COLS; cnt = min(MM_COLS, n - y * MM_COLS); return; }
        y -= nch;
    }
}

__device__ __forceinline__ void ldsm_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t &r0, uint32_t &r1, uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(saddr));
}
// D = A(16x32 s8, row) * B(32x8 s8, col), int32
__device__ __forceinline__ void mma_s8_k32(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}
__device__ __forceinline__ void mma_s8_k16(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};"
                 : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]) : "r"(a0), "r"(a1), "r"(b0), "r"(0));
}
// 32 bits at a 2-byte aligned shared-memory address
__device__ __forceinline__ uint32_t lds_a2(const uint8_t *p) {
    return (uint32_t)*(const unsigned short *)p | ((uint32_t)*(const unsigned short *)(p + 2) << 16);
}
__device__ __forceinline__ float h2f(uint32_t bits) { return __half2float(__ushort_as_half((unsigned short)bits)); }

// B fragments of one 32-k sub-block for the n-tiles 0..NT-1 (activation rows = token columns, MM_ASTRIDE apart)
template <int NT>
__device__ __forceinline__ void load_b(uint32_t aq_saddr, int kbyte, int lane, uint32_t (&b)[NT][2]) {
    const int mi = lane >> 3, rr = lane & 7;
    if (NT == 1) {
        ldsm_x2(b[0][0], b[0][1], aq_saddr + (uint32_t)rr * MM_ASTRIDE + (uint32_t)kbyte + (uint32_t)(mi & 1) * 16);
    } else {
#pragma unroll
        for (int n2 = 0; n2 < NT / 2; n2++)
            ldsm_x4(b[2 * n2][0], b[2 * n2][1], b[2 * n2 + (NT > 1)][0], b[2 * n2 + (NT > 1)][1],
                    aq_saddr + (uint32_t)((2 * n2 + (mi >> 1)) * 8 + rr) * MM_ASTRIDE + (uint32_t)kbyte + (uint32_t)(mi & 1) * 16);
    }
}

// one warp: 16 rows x the super-blocks blk0, blk0 + 2, ... of one stage against NT n-tiles of 8 token columns;
// out[nt] = {row R: cols c, c+1; row R+8: c, c+1}
template <int TYPE, int NT>
__device__ __forceinline__ void unit_compute(const uint8_t *stage, uint32_t rstride, uintptr_t src0, uint32_t rb, int blk0, int nbk, uint32_t aq_saddr,
                                             const float *ad, const int16_t *as32, int lane, float (&out)[NT][4]) {
    const int R = lane >> 2, kq = (lane & 3) * 4, cq = (lane & 3) * 2;
    // every row piece keeps the 16-byte phase of its global address (bulk copies move whole 16-byte lines)
    const uint8_t *row0 = stage + (size_t)R * rstride + ((src0 + (uintptr_t)R * rb) & 15);
    const uint8_t *row1 = stage + (size_t)(R + 8) * rstride + ((src0 + (uintptr_t)(R + 8) * rb) & 15);
#pragma unroll 1
    for (int blk = blk0; blk < nbk; blk += 2) {
        if (TYPE == B200_TYPE_Q4_K || TYPE == B200_TYPE_Q5_K) {
            constexpr int BB = TYPE == B200_TYPE_Q4_K ? 144 : 176;
            const uint8_t *b0p = row0 + blk * BB, *b1p = row1 + blk * BB;
            const uint4 h0 = *(const uint4 *)b0p, h1 = *(const uint4 *)b1p;
            // 6-bit scales / mins of the 8 sub-blocks, packed 4 per word (get_scale_min_k4, ggml-quants.c:631)
            const uint32_t sc0_lo = h0.y & 0x3f3f3f3fu, mn0_lo = h0.z & 0x3f3f3f3fu, sc0_hi = (h0.w & 0x0f0f0f0fu) | ((h0.y >> 2) & 0x30303030u), mn0_hi = ((h0.w >> 4) & 0x0f0f0f0fu) | ((h0.z >> 2) & 0x30303030u);
            const uint32_t sc1_lo = h1.y & 0x3f3f3f3fu, mn1_lo = h1.z & 0x3f3f3f3fu, sc1_hi = (h1.w & 0x0f0f0f0fu) | ((h1.y >> 2) & 0x30303030u), mn1_hi = ((h1.w >> 4) & 0x0f0f0f0fu) | ((h1.z >> 2) & 0x30303030u);
            const uint8_t *q0 = b0p + (TYPE == B200_TYPE_Q5_K ? 48 : 16), *q1 = b1p + (TYPE == B200_TYPE_Q5_K ? 48 : 16);
            uint32_t hb00 = 0, hb01 = 0, hb10 = 0, hb11 = 0;
            if (TYPE == B200_TYPE_Q5_K) { hb00 = *(const uint32_t *)(b0p + 16 + kq); hb01 = *(const uint32_t *)(b0p + 32 + kq); hb10 = *(const uint32_t *)(b1p + 16 + kq); hb11 = *(const uint32_t *)(b1p + 32 + kq); }
            int P[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; nt++) P[nt][0] = P[nt][1] = P[nt][2] = P[nt][3] = 0;
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const uint32_t w00 = *(const uint32_t *)(q0 + 32 * g + kq), w01 = *(const uint32_t *)(q0 + 32 * g + 16 + kq);
                const uint32_t w10 = *(const uint32_t *)(q1 + 32 * g + kq), w11 = *(const uint32_t *)(q1 + 32 * g + 16 + kq);
#pragma unroll
                for (int sub = 0; sub < 2; sub++) {
                    const int j = 2 * g + sub;
                    uint32_t a0 = sub ? (w00 >> 4) & 0x0f0f0f0fu : w00 & 0x0f0f0f0fu, a2 = sub ? (w01 >> 4) & 0x0f0f0f0fu : w01 & 0x0f0f0f0fu;
                    uint32_t a1 = sub ? (w10 >> 4) & 0x0f0f0f0fu : w10 & 0x0f0f0f0fu, a3 = sub ? (w11 >> 4) & 0x0f0f0f0fu : w11 & 0x0f0f0f0fu;
                    if (TYPE == B200_TYPE_Q5_K) {
                        a0 |= ((hb00 >> j) & 0x01010101u) << 4; a2 |= ((hb01 >> j) & 0x01010101u) << 4;
                        a1 |= ((hb10 >> j) & 0x01010101u) << 4; a3 |= ((hb11 >> j) & 0x01010101u) << 4;
                    }
                    const int s0 = (int)(((j < 4 ? sc0_lo : sc0_hi) >> (8 * (j & 3))) & 0xffu), s1 = (int)(((j < 4 ? sc1_lo : sc1_hi) >> (8 * (j & 3))) & 0xffu);
                    uint32_t b[NT][2];
                    load_b<NT>(aq_saddr, blk * 256 + j * 32, lane, b);
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) {
                        int c[4];
                        mma_s8_k32(c, a0, a1, a2, a3, b[nt][0], b[nt][1]);
                        P[nt][0] += s0 * c[0]; P[nt][1] += s0 * c[1]; P[nt][2] += s1 * c[2]; P[nt][3] += s1 * c[3];
                    }
                }
            }
            const float d0 = h2f(h0.x), dm0 = -h2f(h0.x >> 16), d1 = h2f(h1.x), dm1 = -h2f(h1.x >> 16);
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int cc = 0; cc < 2; cc++) {
                    const int col = nt * 8 + cq + cc;
                    const uint4 s = *(const uint4 *)(as32 + ((size_t)col * MM_KB + blk) * 8);
                    const float da = ad[col * MM_KB + blk];
                    int M0 = __dp2a_lo((int)s.x, (int)mn0_lo, 0); M0 = __dp2a_hi((int)s.y, (int)mn0_lo, M0); M0 = __dp2a_lo((int)s.z, (int)mn0_hi, M0); M0 = __dp2a_hi((int)s.w, (int)mn0_hi, M0);
                    int M1 = __dp2a_lo((int)s.x, (int)mn1_lo, 0); M1 = __dp2a_hi((int)s.y, (int)mn1_lo, M1); M1 = __dp2a_lo((int)s.z, (int)mn1_hi, M1); M1 = __dp2a_hi((int)s.w, (int)mn1_hi, M1);
                    out[nt][cc]     = fmaf(da, fmaf(d0, (float)P[nt][cc], dm0 * (float)M0), out[nt][cc]);
                    out[nt][2 + cc] = fmaf(da, fmaf(d1, (float)P[nt][2 + cc], dm1 * (float)M1), out[nt][2 + cc]);
                }
        } else if (TYPE == B200_TYPE_Q6_K) {
            // ql[128] | qh[64] | int8 scales[16] | half d; blocks 2-byte aligned; signed quants q - 32: no offset term
            const uint8_t *b0p = row0 + blk * 210, *b1p = row1 + blk * 210;
            int P[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; nt++) P[nt][0] = P[nt][1] = P[nt][2] = P[nt][3] = 0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t qh00 = lds_a2(b0p + 128 + 32 * h + kq), qh01 = lds_a2(b0p + 128 + 32 * h + 16 + kq);
                const uint32_t qh10 = lds_a2(b1p + 128 + 32 * h + kq), qh11 = lds_a2(b1p + 128 + 32 * h + 16 + kq);
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int j = 4 * h + t;
                    const uint32_t l00 = lds_a2(b0p + 64 * h + 32 * (t & 1) + kq), l01 = lds_a2(b0p + 64 * h + 32 * (t & 1) + 16 + kq);
                    const uint32_t l10 = lds_a2(b1p + 64 * h + 32 * (t & 1) + kq), l11 = lds_a2(b1p + 64 * h + 32 * (t & 1) + 16 + kq);
                    auto q6 = [&](uint32_t l, uint32_t hh) { return __vsub4(((t < 2 ? l : l >> 4) & 0x0f0f0f0fu) | (((hh >> (2 * t)) & 0x03030303u) << 4), 0x20202020u); };
                    const uint32_t a00 = q6(l00, qh00), a01 = q6(l01, qh01), a10 = q6(l10, qh10), a11 = q6(l11, qh11);
                    const int sA0 = (int)(int8_t)b0p[192 + 2 * j], sB0 = (int)(int8_t)b0p[193 + 2 * j], sA1 = (int)(int8_t)b1p[192 + 2 * j], sB1 = (int)(int8_t)b1p[193 + 2 * j];
                    uint32_t b[NT][2];
                    load_b<NT>(aq_saddr, blk * 256 + j * 32, lane, b);
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) {
                        int c[4], e[4];
                        mma_s8_k16(c, a00, a10, b[nt][0]);      // elements 0..15 of the sub-block: scale 2j
                        mma_s8_k16(e, a01, a11, b[nt][1]);      // elements 16..31: scale 2j+1
                        P[nt][0] += sA0 * c[0] + sB0 * e[0]; P[nt][1] += sA0 * c[1] + sB0 * e[1];
                        P[nt][2] += sA1 * c[2] + sB1 * e[2]; P[nt][3] += sA1 * c[3] + sB1 * e[3];
                    }
                }
            }
            const float d0 = h2f(*(const unsigned short *)(b0p + 208)), d1 = h2f(*(const unsigned short *)(b1p + 208));
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int cc = 0; cc < 2; cc++) {
                    const float da = ad[(nt * 8 + cq + cc) * MM_KB + blk];
                    out[nt][cc] = fmaf(d0 * da, (float)P[nt][cc], out[nt][cc]);
                    out[nt][2 + cc] = fmaf(d1 * da, (float)P[nt][2 + cc], out[nt][2 + cc]);
                }
        } else {
            // q4_0 (18 B) / q8_0 (34 B) blocks of 32: one mma per block, one float scale per (row, column, block)
            constexpr int SB = TYPE == B200_TYPE_Q4_0 ? 18 : 34;
#pragma unroll 2
            for (int sbk = 0; sbk < 8; sbk++) {
                const uint8_t *p0 = row0 + (blk * 8 + sbk) * SB, *p1 = row1 + (blk * 8 + sbk) * SB;
                uint32_t a0, a1, a2, a3;
                if (TYPE == B200_TYPE_Q4_0) {
                    const uint32_t w0 = lds_a2(p0 + 2 + kq), w1 = lds_a2(p1 + 2 + kq);
                    a0 = __vsub4(w0 & 0x0f0f0f0fu, 0x08080808u); a2 = __vsub4((w0 >> 4) & 0x0f0f0f0fu, 0x08080808u);
                    a1 = __vsub4(w1 & 0x0f0f0f0fu, 0x08080808u); a3 = __vsub4((w1 >> 4) & 0x0f0f0f0fu, 0x08080808u);
                } else {
                    a0 = lds_a2(p0 + 2 + kq); a2 = lds_a2(p0 + 18 + kq); a1 = lds_a2(p1 + 2 + kq); a3 = lds_a2(p1 + 18 + kq);
                }
                const float d0 = h2f(*(const unsigned short *)p0), d1 = h2f(*(const unsigned short *)p1);
                uint32_t b[NT][2];
                load_b<NT>(aq_saddr, blk * 256 + sbk * 32, lane, b);
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
                    int c[4];
                    mma_s8_k32(c, a0, a1, a2, a3, b[nt][0], b[nt][1]);
#pragma unroll
                    for (int cc = 0; cc < 2; cc++) {
                        const float da = ad[(nt * 8 + cq + cc) * (MM_KB * 8) + blk * 8 + sbk];
                        out[nt][cc] = fmaf((float)c[cc], d0 * da, out[nt][cc]);
                        out[nt][2 + cc] = fmaf((float)c[2 + cc], d1 * da, out[nt][2 + cc]);
                    }
                }
            }
        }
    }
}

template <int TYPE, int NT>
__global__ void __launch_bounds__(MM_THREADS, 1) b200_gemv_mma_kernel(const MmParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = (uint64_t *)smem, *empty = full + MM_MAX_STAGES, *act_full = empty + MM_MAX_STAGES;
    uint8_t *aq = smem + p.off_aq;
    float *ad = (float *)(smem + p.off_ad);
    int16_t *as32 = (int16_t *)(smem + p.off_as);
    float *comb = (float *)(smem + p.off_comb);
    uint8_t *ring = smem + p.off_ring;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ns = p.nstages;
    if (p.use_pdl) pdl_trigger();
    const int r = blockIdx.x % p.nranges, g = blockIdx.x / p.nranges;
    if (g >= p.ngroups) return;
    // column maps: activation column (byte offset into p.act) and destination (element offset into p.dst) of every token column
    __shared__ uint32_t s_colact[MM_COLS];
    __shared__ long long s_coldst[MM_COLS];
    __shared__ int s_grp[2];
    int ncols = p.ncols;
    const uint8_t *Wb = p.W;
    float *partb = p.part;
    if (p.g_E) {
        if (p.use_pdl) pdl_wait();             // the routing tables come from the grouping kernel just before
        if (threadIdx.x == 0) {
            int e, first, cnt;
            group_lookup(p, blockIdx.y, e, first, cnt);
            s_grp[0] = cnt; s_grp[1] = e;
            for (int c = 0; c < cnt; c++) {
                const int pair = p.g_pairs[first + c], t = pair / p.g_n_used, sl = pair % p.g_n_used;
                s_colact[c] = (uint32_t)((p.g_b_ne1 == 1 ? t : pair) * (long long)p.L.col_bytes);
                s_coldst[c] = (long long)sl * (long long)p.g_d_nb1 + (long long)t * (long long)p.g_d_nb2;
            }
        }
        __syncthreads();
        ncols = s_grp[0];
        if (ncols == 0) return;
        Wb += (size_t)s_grp[1] * p.g_expert_stride;
        partb += (size_t)blockIdx.y * p.nranges * MM_COLS * p.N;
    } else {
        if (threadIdx.x < MM_COLS) { s_colact[threadIdx.x] = (uint32_t)(threadIdx.x * p.L.col_bytes); s_coldst[threadIdx.x] = (long long)threadIdx.x * (long long)p.dst_stride; }
        __syncthreads();
    }
    const int t0 = (int)((long long)p.ntiles * g / p.ngroups), t1 = (int)((long long)p.ntiles * (g + 1) / p.ngroups);
    const int nunits = t1 - t0;
    const int nbk = min(MM_KB, p.nb - r * MM_KB);                 // super-blocks in this K-range
    const uint32_t piece = (uint32_t)nbk * p.bbytes;              // bytes of one row in this range

    if (warp == 2 * MM_MAX_STAGES) {
        // ---------------------------------------------------------------- producer: lane l < 16 copies row l of every unit
        if (lane < ns) { mbar_init(&full[lane], 32); mbar_init(&empty[lane], 2); }
        if (lane == 0) mbar_init(act_full, 1);
        mbar_fence_init();
        __syncwarp();
        asm volatile("bar.arrive 1, %0;" ::"r"(MM_THREADS) : "memory");
        const uint64_t pol = p.stream_once ? l2_policy_evict_first() : l2_policy_evict_normal();
        if (p.use_pdl && !p.w_const) pdl_wait();          // src0 written earlier in the same graph (a KV-cache view): no early streaming
        for (int u = 0; u < nunits; u++) {
            const int st = u % ns, use = u / ns;
            if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
            // 16 row pieces of ~1-2 KB: 16-byte cp.async chunks (two lanes per row) completing on the stage's mbarrier.  One bulk copy per
            // piece costs ~60 ns of TMA issue each -- 1 us per 18 KB stage, a 2.7 TB/s ceiling for the whole chip (profiles/r2_ncu_summaries.md)
            {
                const uint8_t *src = Wb + (size_t)((t0 + u) * MM_ROWS + (lane >> 1)) * p.rb + (size_t)r * MM_KB * p.bbytes;
                const uint32_t extra = (uint32_t)((uintptr_t)src & 15);
                const int nch = (int)((extra + piece + 15u) >> 4);
                const uint32_t dsts = smem_u32(ring + (size_t)st * p.stage_bytes + (size_t)(lane >> 1) * p.rstride);
                src -= extra;
                for (int ch = lane & 1; ch < nch; ch += 2)
                    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dsts + ch * 16), "l"(src + ch * 16), "l"(pol) : "memory");
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[st])) : "memory");
            }
            if (u == min(ns, nunits) - 1) {
                // the weight ring is full: now the activations of this K-range (written by the previous kernel), one bulk copy per column
                if (p.use_pdl) pdl_wait();
                if (lane == 0) mbar_arrive_expect_tx(act_full, (uint32_t)ncols * (uint32_t)nbk * 256u);
                __syncwarp();
                if (lane < ncols) bulk_g2s(aq + (size_t)lane * MM_ASTRIDE, p.act + s_colact[lane] + (size_t)r * MM_KR, (uint32_t)nbk * 256u, act_full);
            }
        }
        return;
    }
    // -------------------------------------------------------------------- consumers: scales and sums of this K-range -> shared memory
    if (p.use_pdl) pdl_wait();
    {
        const int q8k = p.L.q8k;
        constexpr int nthr = 2 * MM_MAX_STAGES * 32;
        // token columns beyond ncols (padding of the last n-tile) multiply zeros
        for (int i = threadIdx.x; i < (NT * 8 - ncols) * (MM_KR / 16); i += nthr)
            *(uint4 *)(aq + (size_t)(ncols + i / (MM_KR / 16)) * MM_ASTRIDE + (i % (MM_KR / 16)) * 16) = make_uint4(0, 0, 0, 0);
        if (q8k) {
            // one (column, super-block) per thread: d and the 8 per-32 sums (pairs of the q8_K per-16 bsums)
            for (int i = threadIdx.x; i < NT * 8 * MM_KB; i += nthr) {
                const int col = i / MM_KB, b = i % MM_KB;
                float d = 0.0f;
                uint4 sv = make_uint4(0, 0, 0, 0);
                if (col < ncols && b < nbk) {
                    const uint8_t *cb = p.act + s_colact[col];
                    d = ((const float *)(cb + p.L.off_d))[r * MM_KB + b];
                    const uint4 *bs = (const uint4 *)(cb + p.L.off_sums) + (size_t)(r * MM_KB + b) * 2;
                    const uint4 x = bs[0], y = bs[1];
                    auto pair_sum = [](uint32_t w0, uint32_t w1) {       // {s0, s1}, {s2, s3} (int16 pairs) -> {s0 + s1, s2 + s3}
                        const int a = (int)(short)(w0 & 0xffff) + (int)(short)(w0 >> 16), c = (int)(short)(w1 & 0xffff) + (int)(short)(w1 >> 16);
                        return (uint32_t)(a & 0xffff) | ((uint32_t)c << 16);
                    };
                    sv = make_uint4(pair_sum(x.x, x.y), pair_sum(x.z, x.w), pair_sum(y.x, y.y), pair_sum(y.z, y.w));
                }
                ad[i] = d;
                *(uint4 *)(as32 + (size_t)i * 8) = sv;
            }
        } else {
            for (int i = threadIdx.x; i < NT * 8 * MM_KB * 8; i += nthr) {
                const int col = i / (MM_KB * 8), b = i % (MM_KB * 8);
                ad[i] = (col < ncols && b < nbk * 8) ? ((const float *)(p.act + s_colact[col] + p.L.off_d))[r * MM_KB * 8 + b] : 0.0f;
            }
        }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(MM_THREADS) : "memory");       // scales in place, mbarriers initialised
    const int pair = warp >> 1, half = warp & 1;
    if (pair >= ns) return;
    const uint32_t aq_saddr = smem_u32(aq);
    float *cb = comb + pair * (16 * 32);
    mbar_wait(act_full, 0);
    for (int u = pair, use = 0; u < nunits; u += ns, use++) {
        const uint8_t *src0 = Wb + (size_t)((t0 + u) * MM_ROWS) * p.rb + (size_t)r * MM_KB * p.bbytes;
        const uint8_t *stage = ring + (size_t)pair * p.stage_bytes;
        float out[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; nt++) out[nt][0] = out[nt][1] = out[nt][2] = out[nt][3] = 0.0f;
        mbar_wait(&full[pair], use & 1);
        unit_compute<TYPE, NT>(stage, p.rstride, (uintptr_t)src0, p.rb, half, nbk, aq_saddr, ad, as32, lane, out);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[pair]);
        // ---- the odd-block warp hands its partial to the even-block warp (fixed order: even + odd) ----
        if (half == 1) {
            if (use > 0) asm volatile("bar.sync %0, 64;" ::"r"(9 + pair) : "memory");       // the previous hand-off has been read
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) cb[(nt * 4 + e) * 32 + lane] = out[nt][e];
            asm volatile("bar.sync %0, 64;" ::"r"(2 + pair) : "memory");
            continue;
        }
        asm volatile("bar.sync %0, 64;" ::"r"(2 + pair) : "memory");
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int e = 0; e < 4; e++) out[nt][e] += cb[(nt * 4 + e) * 32 + lane];
        if (u + ns < nunits) asm volatile("bar.arrive %0, 64;" ::"r"(9 + pair) : "memory");
        // ---- store: rows R, R+8 of the tile; columns nt * 8 + cq + {0, 1} ----
        const int R = lane >> 2, cq = (lane & 3) * 2, row = (t0 + u) * MM_ROWS + R;
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int cc = 0; cc < 2; cc++) {
                const int col = nt * 8 + cq + cc;
                if (col < ncols) {
                    if (p.nranges > 1) {
                        float *pp = partb + ((size_t)r * MM_COLS + col) * p.N;
                        pp[row] = out[nt][cc]; pp[row + 8] = out[nt][2 + cc];
                    } else {
                        float *dp = p.dst + s_coldst[col];
                        float r0 = 0.0f, r1 = 0.0f;
                        if (p.residual) { const float *rp = p.residual + s_coldst[col]; r0 = rp[row]; r1 = rp[row + 8]; }
                        dp[row] = out[nt][cc] + r0; dp[row + 8] = out[nt][2 + cc] + r1;
                    }
                }
            }
    }
}

// dst[col][row] = sum over the K-ranges in range order (deterministic)
__global__ void __launch_bounds__(256) b200_gemv_mma_reduce_kernel(const float *__restrict__ part, int nranges, int N, int ncols, float *dst, size_t dst_stride,
                                                                   const float *residual, int use_pdl) {
    if (use_pdl) { pdl_trigger(); pdl_wait(); }
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)ncols * N) return;
    const int col = (int)(i / N), row = (int)(i % N);
    float a = part[(size_t)col * N + row];
    for (int r = 1; r < nranges; r++) a += part[((size_t)r * MM_COLS + col) * N + row];
    if (residual) a += residual[(size_t)col * dst_stride + row];
    dst[(size_t)col * dst_stride + row] = a;
}

// grouped: grid = (N / 256, 32 columns, chunks); the chunk's pairs give the destination columns
__global__ void __launch_bounds__(256) b200_gemv_mma_reduce_grouped_kernel(const MmParams p) {
    if (p.use_pdl) { pdl_trigger(); pdl_wait(); }
    int e, first, cnt;
    group_lookup(p, blockIdx.z, e, first, cnt);
    const int col = blockIdx.y, row = blockIdx.x * 256 + threadIdx.x;
    if (col >= cnt || row >= p.N) return;
    const int pair = p.g_pairs[first + col], t = pair / p.g_n_used, sl = pair % p.g_n_used;
    const float *pp = p.part + ((size_t)blockIdx.z * p.nranges * MM_COLS + col) * p.N + row;
    float a = pp[0];
    for (int r = 1; r < p.nranges; r++) a += pp[(size_t)r * MM_COLS * p.N];
    p.dst[(size_t)sl * p.g_d_nb1 + (size_t)t * p.g_d_nb2 + row] = a;
}

template <int TYPE, int NT>
int launch_t(b200_ctx *ctx, const MmParams &p, int grid, size_t smem, int grid_y = 1) {
    auto kern = b200_gemv_mma_kernel<TYPE, NT>;
    static bool attr_set[16] = {false};
    if (!attr_set[ctx->device & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin - 512));     // 512: the static column maps
        attr_set[ctx->device & 15] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, (unsigned)grid_y); cfg.blockDim = dim3(MM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = p.use_pdl ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
    ctx->launches++;
    return B200_OK;
}
template <int TYPE>
int launch_type(b200_ctx *ctx, const MmParams &p, int grid, size_t smem, int grid_y = 1) {
    if (p.g_E) return launch_t<TYPE, 4>(ctx, p, grid, smem, grid_y);
    return p.ncols <= 8 ? launch_t<TYPE, 1>(ctx, p, grid, smem) : p.ncols <= 16 ? launch_t<TYPE, 2>(ctx, p, grid, smem) : launch_t<TYPE, 4>(ctx, p, grid, smem);
}

}  // namespace

bool gemv_mma_supported(int type, int64_t N, int64_t K, int64_t M) {
    if (!b200_type_is_quant(type) || getenv("GGML_B200_NO_GEMV_MMA")) return false;
    return N % MM_ROWS == 0 && K % 256 == 0 && M >= 1 && N >= MM_ROWS;
}

// dst[col * dst_stride + n] = W[n, :] . x[:, col] for `ncols` <= 32 columns pre-quantised in `act` (ActLayout scratch)
static int mma_run(b200_ctx *ctx, MmParams &p, int type, int64_t N, int64_t K, int grid_y) {
    p.type = type; p.N = (int)N; p.K = (int)K;
    p.L = ActLayout::make(b200_act_mode_q8k(type), K);
    p.nb = (int)(K / 256);
    p.bbytes = type == B200_TYPE_Q4_K ? 144u : type == B200_TYPE_Q5_K ? 176u : type == B200_TYPE_Q6_K ? 210u : type == B200_TYPE_Q4_0 ? 144u : 272u;
    p.nranges = (p.nb + MM_KB - 1) / MM_KB;
    p.ntiles = (int)(N / MM_ROWS);
    // dense: the machine is one wave of (tile group, K range) CTAs; grouped: the chunks of all experts share it
    int G = ctx->sm_count;
    if (grid_y > 1) { G = (4 * ctx->sm_count + grid_y - 1) / grid_y; if (G > ctx->sm_count) G = ctx->sm_count; if (G < p.nranges) G = p.nranges; }
    p.ngroups = G / p.nranges;
    if (p.ngroups < 1) p.ngroups = 1;
    if (p.ngroups > p.ntiles) p.ngroups = p.ntiles;
    p.use_pdl = ctx->opt_pdl;
    // shared memory: barriers | activation quants [32][2048 + 16] | scales | per-32 sums | pair hand-off | ring
    const int q8k = p.L.q8k;
    uint32_t off = (2 * MM_MAX_STAGES + 1) * 8;
    off = (off + 127) & ~127u;
    p.off_aq = off; off += MM_COLS * MM_ASTRIDE;
    p.off_ad = off; off += MM_COLS * MM_KB * (q8k ? 1 : 8) * 4;
    p.off_as = off; off += q8k ? MM_COLS * MM_KB * 8 * 2 : 0;
    p.off_comb = off; off += MM_MAX_STAGES * 16 * 32 * 4;
    off = (off + 127) & ~127u;
    p.off_ring = off;
    p.rstride = ((MM_KB * p.bbytes + 15u) & ~15u) + 16;           // + worst-case misalignment of a row start
    p.stage_bytes = (MM_ROWS * p.rstride + 127u) & ~127u;
    int ns = (int)((ctx->smem_optin - 512 - off) / p.stage_bytes);
    if (ns > MM_MAX_STAGES) ns = MM_MAX_STAGES;
    if (ns < 2) { b200_set_error("gemv_mma: shared memory"); return B200_ERR_FAILED; }
    p.nstages = ns;
    const size_t smem = (size_t)off + (size_t)ns * p.stage_bytes;
    if (p.nranges > 1) {
        p.part = (float *)ctx->get_scratch(SCRATCH_MISC, (size_t)grid_y * p.nranges * MM_COLS * N * 4);
        if (!p.part) return B200_ERR_ALLOC;
    }
    const int grid = p.ngroups * p.nranges;
    switch (type) {
        case B200_TYPE_Q4_K: return launch_type<B200_TYPE_Q4_K>(ctx, p, grid, smem, grid_y);
        case B200_TYPE_Q5_K: return launch_type<B200_TYPE_Q5_K>(ctx, p, grid, smem, grid_y);
        case B200_TYPE_Q6_K: return launch_type<B200_TYPE_Q6_K>(ctx, p, grid, smem, grid_y);
        case B200_TYPE_Q4_0: return launch_type<B200_TYPE_Q4_0>(ctx, p, grid, smem, grid_y);
        default:             return launch_type<B200_TYPE_Q8_0>(ctx, p, grid, smem, grid_y);
    }
}

int launch_gemv_mma(b200_ctx *ctx, int type, const uint8_t *W, size_t rb, int64_t N, int64_t K, const uint8_t *act, int ncols, float *dst, size_t dst_stride, bool stream_once, bool w_const, const float *residual) {
    MmParams p = {};
    p.residual = residual;
    p.stream_once = stream_once ? 1 : 0; p.w_const = w_const ? 1 : 0;
    p.W = W; p.rb = (uint32_t)rb; p.ncols = ncols;
    p.act = act;
    p.dst = dst; p.dst_stride = dst_stride;
    int rc = mma_run(ctx, p, type, N, K, 1);
    if (rc || p.nranges == 1) return rc;
    const int64_t total = (int64_t)ncols * N;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((total + 255) / 256)); cfg.blockDim = dim3(256); cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = p.use_pdl ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, b200_gemv_mma_reduce_kernel, (const float *)p.part, p.nranges, (int)N, ncols, dst, dst_stride, residual, p.use_pdl));
    ctx->launches++;
    return B200_OK;
}

// MUL_MAT_ID over pairs grouped by expert on the device: dst[:, slot, tok] = W[expert(pair)] . act[:, column(pair)]; every chunk of
// <= 32 pairs of one expert streams that expert's matrix once.  No host knowledge of the routing: the grid covers the worst case
// (g.max_chunks) and chunks beyond the real count exit at once.
int launch_gemv_mma_grouped(b200_ctx *ctx, int type, const uint8_t *W, size_t rb, int64_t N, int64_t K, const uint8_t *act, const MmGroupDesc &g, float *dst) {
    MmParams p = {};
    p.stream_once = 0; p.w_const = 1;
    p.W = W; p.rb = (uint32_t)rb; p.ncols = MM_COLS;
    p.act = act; p.dst = dst;
    p.g_off = g.off; p.g_pairs = g.pairs; p.g_E = g.E; p.g_n_used = g.n_used; p.g_b_ne1 = g.b_ne1;
    p.g_expert_stride = g.expert_stride; p.g_d_nb1 = g.d_nb1; p.g_d_nb2 = g.d_nb2;
    int rc = mma_run(ctx, p, type, N, K, g.max_chunks);
    if (rc || p.nranges == 1) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((N + 255) / 256), MM_COLS, (unsigned)g.max_chunks); cfg.blockDim = dim3(256); cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = p.use_pdl ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, b200_gemv_mma_reduce_grouped_kernel, p));
    ctx->launches++;
    return B200_OK;
}
