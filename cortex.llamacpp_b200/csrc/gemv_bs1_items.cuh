// gemv_bs1_items.cuh -- block decoders of the batch-1 K-quant GEMV (gemv_bs1.cu), host/device portable like gemv_items.cuh so
// that tests/host_emul can run the very same bit-twiddling on the CPU against the oracle's exact integer block sums
// (ggml_vec_dot_q{4,5,6}_K_q8_K, ggml-cpu-quants.c:6511, 7326, 8148; scalar forms :7267-7322).
//   block_q45k<Q5> : one 256-weight Q4_K / Q5_K block against 256 int8 activations (272-byte padded) + 8 per-32 sums
//   item_q6k       : one 128-weight half of a Q6_K block (2-byte aligned) against 128 int8 activations (144-byte padded) + 8 per-16 sums
#pragma once
#include "gemv_items.cuh"

namespace bs1 {

using gemv::U4; using gemv::ld128; using gemv::dp4a_ss; using gemv::h2f_bits;

B200_HD int dp4a_us(uint32_t a, uint32_t b, int c) {      // unsigned bytes x signed bytes
#ifdef __CUDA_ARCH__
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    for (int i = 0; i < 4; i++) c += (int)(uint8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i));
    return c;
#endif
}
B200_HD int dp2a_lo_su(uint32_t a16x2, uint32_t b8, int c) {   // a.lo16*b.byte0 + a.hi16*b.byte1 (signed 16 x unsigned 8)
#ifdef __CUDA_ARCH__
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16x2), "r"(b8), "r"(c));
    return d;
#else
    return c + (int)(int16_t)(a16x2 & 0xffffu) * (int)(b8 & 0xffu) + (int)(int16_t)(a16x2 >> 16) * (int)((b8 >> 8) & 0xffu);
#endif
}
B200_HD int dp2a_hi_su(uint32_t a16x2, uint32_t b8, int c) {   // a.lo16*b.byte2 + a.hi16*b.byte3
#ifdef __CUDA_ARCH__
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16x2), "r"(b8), "r"(c));
    return d;
#else
    return c + (int)(int16_t)(a16x2 & 0xffffu) * (int)((b8 >> 16) & 0xffu) + (int)(int16_t)(a16x2 >> 16) * (int)(b8 >> 24);
#endif
}
B200_HD uint32_t byte_perm(uint32_t lo, uint32_t hi, uint32_t sel) {     // PRMT, default mode (selectors 0..7)
#ifdef __CUDA_ARCH__
    return __byte_perm(lo, hi, sel);
#else
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
#endif
}

// Whole 256-weight block per lane (16 lanes walk one row, the two half-warps take two rows of the stage at once): the 16-byte
// header, the 6-bit scale unpack, d/dmin and the activation scale are paid once per 256 weights instead of once per 64
// (~54 instead of ~80 instructions per 64 weights), and one shuffle tree reduces two rows.
// Activations: aq at 272 bytes per block (256 + 16 pad: conflict-free 128-bit loads at lane stride), s32 8 x int16 per block.
template <bool Q5>
B200_HD float block_q45k(const uint8_t *b, const uint8_t *ap, U4 sums, float da, int *dbgP = nullptr, int *dbgM = nullptr) {
    const U4 hdr = ld128(b);
    const uint32_t sc_lo = hdr.y & 0x3f3f3f3fu, mn_lo = hdr.z & 0x3f3f3f3fu;
    const uint32_t sc_hi = (hdr.w & 0x0f0f0f0fu) | ((hdr.y >> 2) & 0x30303030u);
    const uint32_t mn_hi = ((hdr.w >> 4) & 0x0f0f0f0fu) | ((hdr.z >> 2) & 0x30303030u);
    U4 ha, hb;
    if (Q5) { ha = ld128(b + 16); hb = ld128(b + 32); }
    int P = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint8_t *qp = b + (Q5 ? 48 : 16) + g * 32;
        const U4 qa = ld128(qp), qb = ld128(qp + 16);
        const U4 a0 = ld128(ap + g * 64), a1 = ld128(ap + g * 64 + 16);
        const U4 a2 = ld128(ap + g * 64 + 32), a3 = ld128(ap + g * 64 + 48);
        int sA, sB;
        if (!Q5) {
            int s0 = dp4a_ss((uint32_t)(qa.x & 0x0f0f0f0fu), a0.x, 0), s1 = dp4a_ss((uint32_t)(qa.y & 0x0f0f0f0fu), a0.y, 0);
            s0 = dp4a_ss((uint32_t)(qa.z & 0x0f0f0f0fu), a0.z, s0); s1 = dp4a_ss((uint32_t)(qa.w & 0x0f0f0f0fu), a0.w, s1);
            s0 = dp4a_ss((uint32_t)(qb.x & 0x0f0f0f0fu), a1.x, s0); s1 = dp4a_ss((uint32_t)(qb.y & 0x0f0f0f0fu), a1.y, s1);
            s0 = dp4a_ss((uint32_t)(qb.z & 0x0f0f0f0fu), a1.z, s0); s1 = dp4a_ss((uint32_t)(qb.w & 0x0f0f0f0fu), a1.w, s1);
            sA = s0 + s1;
            int t0 = dp4a_us(qa.x & 0xf0f0f0f0u, a2.x, 0), t1 = dp4a_us(qa.y & 0xf0f0f0f0u, a2.y, 0);      // high nibbles in place: 16x, exact
            t0 = dp4a_us(qa.z & 0xf0f0f0f0u, a2.z, t0); t1 = dp4a_us(qa.w & 0xf0f0f0f0u, a2.w, t1);
            t0 = dp4a_us(qb.x & 0xf0f0f0f0u, a3.x, t0); t1 = dp4a_us(qb.y & 0xf0f0f0f0u, a3.y, t1);
            t0 = dp4a_us(qb.z & 0xf0f0f0f0u, a3.z, t0); t1 = dp4a_us(qb.w & 0xf0f0f0f0u, a3.w, t1);
            sB = (t0 + t1) >> 4;
        } else {
            const uint32_t qw[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            const uint32_t hw[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
            const uint32_t al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const uint32_t ah[8] = {a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
            int s0 = 0, s1 = 0, t0 = 0, t1 = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t hsh = hw[i] >> (2 * g);
                const uint32_t lo = (qw[i] & 0x0f0f0f0fu) | ((hsh << 4) & 0x10101010u);
                const uint32_t hi = ((qw[i] >> 4) & 0x0f0f0f0fu) | ((hsh << 3) & 0x10101010u);
                if (i & 1) { s1 = dp4a_ss((uint32_t)lo, al[i], s1); t1 = dp4a_ss((uint32_t)hi, ah[i], t1); }
                else       { s0 = dp4a_ss((uint32_t)lo, al[i], s0); t0 = dp4a_ss((uint32_t)hi, ah[i], t0); }
            }
            sA = s0 + s1; sB = t0 + t1;
        }
        const uint32_t scw = g < 2 ? sc_lo : sc_hi;
        const int sh = (g & 1) * 16;
        P += (int)((scw >> sh) & 0xffu) * sA + (int)((scw >> (sh + 8)) & 0xffu) * sB;
    }
    int M = dp2a_lo_su(sums.x, mn_lo, 0);
    M = dp2a_hi_su(sums.y, mn_lo, M);
    M = dp2a_lo_su(sums.z, mn_hi, M);
    M = dp2a_hi_su(sums.w, mn_hi, M);
    if (dbgP) { *dbgP = P; *dbgM = M; }
    const float d = h2f_bits(hdr.x), dmin = h2f_bits(hdr.x >> 16);
    return (d * da) * (float)P - (dmin * da) * (float)M;
}

// ---------------------------------------------------------------------------------------------- Q6_K
// item = 128 weights (half h of a 210-byte block: ql 64 B at 64h, qh 32 B at 128+32h, int8 scales 8 B at 192+8h, half d at 208).
// Blocks are only 2-byte aligned: every 4-byte word is fetched as two aligned words and a PRMT whose selector is computed
// from the address, so both alignments run the same instructions.
B200_HD void ld4_words(const uint32_t *W, int word, uint32_t sel, uint32_t (&out)[4]) {
    const uint32_t r0 = W[word], r1 = W[word + 1], r2 = W[word + 2], r3 = W[word + 3], r4 = W[word + 4];
    out[0] = byte_perm(r0, r1, sel); out[1] = byte_perm(r1, r2, sel); out[2] = byte_perm(r2, r3, sel); out[3] = byte_perm(r3, r4, sel);
}
B200_HD float item_q6k(const uint8_t *row, int it, const uint8_t *aq, const U4 *s16, const float *ad, int *dbgP = nullptr) {
    const uint8_t *b = row + (it >> 1) * 210;
    const int h = it & 1;
    const uint32_t mis = (uint32_t)(uintptr_t)b & 2u;
    const uint32_t *W = (const uint32_t *)(b - mis);
    const uint32_t sel = mis ? 0x5432u : 0x3210u;
    const uint8_t *ap = aq + it * 144;
    const U4 bs = s16[it];
    const uint32_t bsw[4] = {bs.x, bs.y, bs.z, bs.w};
    uint32_t S[2];
    {
        const int w = 48 + 2 * h;
        const uint32_t r0 = W[w], r1 = W[w + 1], r2 = W[w + 2];
        S[0] = byte_perm(r0, r1, sel); S[1] = byte_perm(r1, r2, sel);
    }
    int P = 0;
#pragma unroll
    for (int hs = 0; hs < 2; hs++) {
        uint32_t H[4];
        ld4_words(W, 32 + 8 * h + 4 * hs, sel, H);
#pragma unroll
        for (int tl = 0; tl < 2; tl++) {
            uint32_t L[4];
            ld4_words(W, 16 * h + 8 * tl + 4 * hs, sel, L);
#pragma unroll
            for (int nib = 0; nib < 2; nib++) {
                const int t = tl + 2 * nib, sg = 2 * t + hs;
                const U4 a = ld128(ap + 16 * sg);
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
                int s = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t lo = nib ? (L[k] >> 4) : L[k];
                    const uint32_t hi = t == 0 ? (H[k] << 4) : t == 1 ? (H[k] << 2) : t == 2 ? H[k] : (H[k] >> 2);
                    s = dp4a_ss((uint32_t)((lo & 0x0f0f0f0fu) | (hi & 0x30303030u)), aw[k], s);
                }
                // quants kept unsigned (0..63); the -32 offset goes through the activation bsums (exact)
                const int bsum = (int)(int16_t)((sg & 1) ? (bsw[sg >> 1] >> 16) : (bsw[sg >> 1] & 0xffffu));
                const int scale = (int)(int8_t)((S[sg >> 2] >> (8 * (sg & 3))) & 0xffu);
                P += scale * (s - 32 * bsum);
            }
        }
    }
    if (dbgP) *dbgP = P;
    const float d = h2f_bits(*(const uint16_t *)(b + 208));
    return (d * ad[it >> 1]) * (float)P;
}


}  // namespace bs1
