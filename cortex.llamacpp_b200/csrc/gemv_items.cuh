// gemv_items.cuh -- per-lane "item" decoders for the decode GEMV (and reused by the debug block-sum hook).
//
// An ITEM is the unit one lane processes per iteration: a slice of one weight row that sits in shared
// memory exactly as it sits in the GGUF file (block layouts: ggml-common.h:161-328), dotted against the
// int8 activations of the same K-range.  All integer arithmetic reproduces the CPU oracle's per-block
// sums exactly:
//     Q4_K / Q5_K : item = 64 elems  (one 32-byte qs group = sub-blocks 2g, 2g+1)   vs q8_K   (ggml-cpu-quants.c:7267, :7326)
//     Q6_K        : item = 128 elems (one half block)                               vs q8_K   (:8148)
//     Q4_0 / Q8_0 : item = 128 elems (4 blocks)                                     vs q8_0   (:1912, :3663)
// Replaces vec_dot_q*_q8_1 (ggml-cuda/vecdotq.cuh:527-787): no q8_1, 128-bit shared loads where the
// format allows, exact-integer min/offset handling through the activation block sums.
//
// The file is host/device portable (B200_HD) so tests/host_emul can run the very same decode code on
// the CPU against the oracle.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace gemv {

// ---- portable intrinsics --------------------------------------------------------------------------
B200_HD int dp4a_ss(uint32_t a, uint32_t b, int c) {   // signed x signed bytes
#ifdef __CUDA_ARCH__
    return __dp4a((int)a, (int)b, c);
#else
    for (int i = 0; i < 4; i++) c += (int)(int8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i));
    return c;
#endif
}
B200_HD uint32_t align2_word(uint32_t lo, uint32_t hi, int sub2) {   // 32 bits starting at byte sub2 (0 or 2) of lo:hi
#ifdef __CUDA_ARCH__
    return sub2 ? __byte_perm(lo, hi, 0x5432) : lo;
#else
    return sub2 ? ((lo >> 16) | (hi << 16)) : lo;
#endif
}
B200_HD float h2f_bits(uint32_t h) {
#ifdef __CUDA_ARCH__
    return __half2float(__ushort_as_half((unsigned short)(h & 0xffffu)));
#else
    const uint32_t sign = (h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else { int e = -1; do { e++; man <<= 1; } while (!(man & 0x400u)); bits = sign | (uint32_t)(112 - e) << 23 | (man & 0x3ffu) << 13; }
    } else if (exp == 31) bits = sign | 0x7f800000u | man << 13;
    else bits = sign | (exp + 112u) << 23 | man << 13;
    union { uint32_t u; float f; } cv; cv.u = bits; return cv.f;
#endif
}

struct U4 { uint32_t x, y, z, w; };
B200_HD U4 ld128(const void *p) {
#ifdef __CUDA_ARCH__
    const uint4 v = *(const uint4 *)p; U4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
    U4 r; const uint32_t *q = (const uint32_t *)p; r.x = q[0]; r.y = q[1]; r.z = q[2]; r.w = q[3]; return r;
#endif
}
B200_HD uint32_t ld32(const void *p) { return *(const uint32_t *)p; }

// ---- shared-memory view of the quantised activations ------------------------------------------------
// qs are stored per item with 16 bytes of padding (bank-conflict-free 128-bit loads at lane stride),
// column c at q + c*q_stride etc.
struct ActView {
    const int8_t * q;      // [ncols][nitems * (ITEM+16)]
    const float *  d;      // [ncols][K/G]
    const int16_t *s;      // [ncols][K/S]
    uint32_t q_stride, d_stride, s_stride;   // per-column strides in bytes / floats / int16
};

enum { T_Q4_0 = 2, T_Q8_0 = 8, T_Q4_K = 12, T_Q5_K = 13, T_Q6_K = 14 };

template <int TYPE> struct Traits;
template <> struct Traits<T_Q4_K> { static constexpr int ITEM = 64,  BLOCK = 256, BYTES = 144, Q8K = 1; };
template <> struct Traits<T_Q5_K> { static constexpr int ITEM = 64,  BLOCK = 256, BYTES = 176, Q8K = 1; };
template <> struct Traits<T_Q6_K> { static constexpr int ITEM = 128, BLOCK = 256, BYTES = 210, Q8K = 1; };
template <> struct Traits<T_Q4_0> { static constexpr int ITEM = 128, BLOCK = 32,  BYTES = 18,  Q8K = 0; };
template <> struct Traits<T_Q8_0> { static constexpr int ITEM = 128, BLOCK = 32,  BYTES = 34,  Q8K = 0; };

template <int TYPE> B200_HD constexpr int act_item_stride() { return Traits<TYPE>::ITEM + 16; }

// six-bit (scale, min) of sub-block j from the 12 packed bytes (get_scale_min_k4, ggml-quants.c:631-639)
B200_HD void scale_min_k4(uint32_t w0, uint32_t w1, uint32_t w2, int j, int &sc, int &mn) {
    if (j < 4) {
        sc = (w0 >> (8 * j)) & 63;
        mn = (w1 >> (8 * j)) & 63;
    } else {
        const int i = j - 4;
        const uint32_t b2 = (w2 >> (8 * i)) & 0xff;
        sc = (int)((b2 & 0x0f) | ((((w0 >> (8 * i)) & 0xff) >> 6) << 4));
        mn = (int)((b2 >> 4)   | ((((w1 >> (8 * i)) & 0xff) >> 6) << 4));
    }
}

// Result of one item against one column: value to add to the row accumulator, plus (debug) the exact
// integer pieces and which weight block they belong to.
struct DbgSink {
    int32_t *P, *M;     // per-block accumulators of the current row (atomic adds on device)
};

#ifdef __CUDA_ARCH__
#define B200_DBG_ADD(ptr, v) atomicAdd((int *)(ptr), (int)(v))
#else
#define B200_DBG_ADD(ptr, v) (*(ptr) += (v))
#endif

// ---------------------------------------------------------------------------------------------- Q4_K / Q5_K
template <int TYPE, int NC, bool DBG>
B200_HD void dot_item_q45k(const uint8_t *rowp, int it, const ActView &A, float *acc, DbgSink dbg) {
    constexpr int BYTES = Traits<TYPE>::BYTES;
    const uint8_t *b = rowp + (size_t)(it >> 2) * BYTES;
    const int g = it & 3;
    const U4 hdr = ld128(b);
    const float d = h2f_bits(hdr.x), dmin = h2f_bits(hdr.x >> 16);
    int sc0, m0, sc1, m1;
    scale_min_k4(hdr.y, hdr.z, hdr.w, 2 * g, sc0, m0);
    scale_min_k4(hdr.y, hdr.z, hdr.w, 2 * g + 1, sc1, m1);
    const int qs_off = (TYPE == T_Q5_K ? 48 : 16) + g * 32;
    const U4 qa = ld128(b + qs_off), qb = ld128(b + qs_off + 16);
    uint32_t lo[8], hi[8];
    const uint32_t qw[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
    for (int i = 0; i < 8; i++) { lo[i] = qw[i] & 0x0f0f0f0fu; hi[i] = (qw[i] >> 4) & 0x0f0f0f0fu; }
    if (TYPE == T_Q5_K) {
        const U4 ha = ld128(b + 16), hb = ld128(b + 32);
        const uint32_t hw[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            lo[i] |= ((hw[i] >> (2 * g)) & 0x01010101u) << 4;
            hi[i] |= ((hw[i] >> (2 * g + 1)) & 0x01010101u) << 4;
        }
    }
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const int8_t *ap = A.q + (size_t)c * A.q_stride + (size_t)it * act_item_stride<TYPE>();
        const U4 a0 = ld128(ap), a1 = ld128(ap + 16), a2 = ld128(ap + 32), a3 = ld128(ap + 48);
        int sA = 0, sB = 0;
        sA = dp4a_ss(lo[0], a0.x, sA); sA = dp4a_ss(lo[1], a0.y, sA); sA = dp4a_ss(lo[2], a0.z, sA); sA = dp4a_ss(lo[3], a0.w, sA);
        sA = dp4a_ss(lo[4], a1.x, sA); sA = dp4a_ss(lo[5], a1.y, sA); sA = dp4a_ss(lo[6], a1.z, sA); sA = dp4a_ss(lo[7], a1.w, sA);
        sB = dp4a_ss(hi[0], a2.x, sB); sB = dp4a_ss(hi[1], a2.y, sB); sB = dp4a_ss(hi[2], a2.z, sB); sB = dp4a_ss(hi[3], a2.w, sB);
        sB = dp4a_ss(hi[4], a3.x, sB); sB = dp4a_ss(hi[5], a3.y, sB); sB = dp4a_ss(hi[6], a3.z, sB); sB = dp4a_ss(hi[7], a3.w, sB);
        const int P = sc0 * sA + sc1 * sB;
        const int16_t *sp = A.s + (size_t)c * A.s_stride + (size_t)it * 4;
        const uint32_t s01 = ld32(sp), s23 = ld32(sp + 2);
        const int M = m0 * ((int)(int16_t)(s01 & 0xffff) + (int)(int16_t)(s01 >> 16)) +
                      m1 * ((int)(int16_t)(s23 & 0xffff) + (int)(int16_t)(s23 >> 16));
        const float da = A.d[(size_t)c * A.d_stride + (it >> 2)];
        acc[c] += (d * da) * (float)P - (dmin * da) * (float)M;
        if (DBG && c == 0) { B200_DBG_ADD(dbg.P + (it >> 2), P); B200_DBG_ADD(dbg.M + (it >> 2), M); }
    }
}

// ---------------------------------------------------------------------------------------------- 2-byte aligned streams
// word k (32 bits at byte offset 4k) of a stream that starts at the 2-byte aligned address p
B200_HD uint32_t word_at(const uint8_t *p, int k) {
    const uintptr_t a = (uintptr_t)p + 4 * (uintptr_t)k;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    return (a & 2) ? align2_word(w[0], w[1], 2) : w[0];
}
B200_HD uint32_t half_at(const uint8_t *p) { return *(const uint16_t *)p; }

// loads n words starting at the 2-byte aligned p with n+1 aligned 32-bit loads
template <int N> B200_HD void load_words(const uint8_t *p, uint32_t *out) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    if (a & 2) {
        uint32_t prev = w[0];
#pragma unroll
        for (int i = 0; i < N; i++) { const uint32_t nx = w[i + 1]; out[i] = align2_word(prev, nx, 2); prev = nx; }
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) out[i] = w[i];
    }
}

// ---------------------------------------------------------------------------------------------- Q6_K
template <int NC, bool DBG>
B200_HD void dot_item_q6k(const uint8_t *rowp, int it, const ActView &A, float *acc, DbgSink dbg) {
    const uint8_t *b = rowp + (size_t)(it >> 1) * 210;
    const int h = it & 1;
    uint32_t L[16], H[8], S[2];
    load_words<16>(b + 64 * h, L);
    load_words<8>(b + 128 + 32 * h, H);
    load_words<2>(b + 192 + 8 * h, S);
    const float d = h2f_bits(half_at(b + 208));
    // 6-bit quants (unsigned, 0..63) of the 8 scale groups: group sg = 2t + (i>>2), word i of quadrant t
    uint32_t q[32];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        q[i]      = (L[i] & 0x0f0f0f0fu)            | ((H[i] << 4) & 0x30303030u);
        q[8 + i]  = (L[8 + i] & 0x0f0f0f0fu)        | ((H[i] << 2) & 0x30303030u);
        q[16 + i] = ((L[i] >> 4) & 0x0f0f0f0fu)     | (H[i] & 0x30303030u);
        q[24 + i] = ((L[8 + i] >> 4) & 0x0f0f0f0fu) | ((H[i] >> 2) & 0x30303030u);
    }
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const int8_t *ap = A.q + (size_t)c * A.q_stride + (size_t)it * act_item_stride<T_Q6_K>();
        const int16_t *sp = A.s + (size_t)c * A.s_stride + (size_t)it * 8;
        const U4 bs0 = ld128(sp);
        const uint32_t bsw[4] = {bs0.x, bs0.y, bs0.z, bs0.w};
        int P = 0;
#pragma unroll
        for (int sg = 0; sg < 8; sg++) {
            const U4 a = ld128(ap + 16 * sg);
            int s = 0;
            s = dp4a_ss(q[4 * sg + 0], a.x, s); s = dp4a_ss(q[4 * sg + 1], a.y, s);
            s = dp4a_ss(q[4 * sg + 2], a.z, s); s = dp4a_ss(q[4 * sg + 3], a.w, s);
            const int bsum = (int)(int16_t)((sg & 1) ? (bsw[sg >> 1] >> 16) : (bsw[sg >> 1] & 0xffff));
            const int scale = (int)(int8_t)((S[sg >> 2] >> (8 * (sg & 3))) & 0xff);
            P += scale * (s - 32 * bsum);
        }
        const float da = A.d[(size_t)c * A.d_stride + (it >> 1)];
        acc[c] += (d * da) * (float)P;
        if (DBG && c == 0) B200_DBG_ADD(dbg.P + (it >> 1), P);
    }
}

// ---------------------------------------------------------------------------------------------- Q4_0 / Q8_0
// item = up to 4 consecutive 32-element blocks; nvalid = number of blocks of this item inside K
template <int NC, bool DBG>
B200_HD void dot_item_q40(const uint8_t *rowp, int it, int nvalid, const ActView &A, float *acc, DbgSink dbg) {
    const uint8_t *p = rowp + (size_t)it * 72;
    uint32_t W[18];
    load_words<18>(p, W);    // stream of 72 bytes; block bi: half d at 18bi, qs at 18bi+2
#pragma unroll
    for (int bi = 0; bi < 4; bi++) {
        if (bi < nvalid) {
            // bytes 18bi .. 18bi+17 ; word index (18bi)>>2, sub = (18bi)&2
            constexpr int dummy = 0; (void)dummy;
            const int o = 18 * bi;
            const uint32_t dbits = (o & 2) ? (W[o >> 2] >> 16) : (W[o >> 2] & 0xffff);
            uint32_t qw[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int t = o + 2 + 4 * k;
                qw[k] = (t & 2) ? align2_word(W[t >> 2], W[(t >> 2) + 1 < 18 ? (t >> 2) + 1 : 17], 2) : W[t >> 2];
            }
            const float dw = h2f_bits(dbits);
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int8_t *ap = A.q + (size_t)c * A.q_stride + (size_t)it * act_item_stride<T_Q4_0>() + 32 * bi;
                const U4 a0 = ld128(ap), a1 = ld128(ap + 16);
                int s = 0;
                s = dp4a_ss(qw[0] & 0x0f0f0f0fu, a0.x, s); s = dp4a_ss(qw[1] & 0x0f0f0f0fu, a0.y, s);
                s = dp4a_ss(qw[2] & 0x0f0f0f0fu, a0.z, s); s = dp4a_ss(qw[3] & 0x0f0f0f0fu, a0.w, s);
                s = dp4a_ss((qw[0] >> 4) & 0x0f0f0f0fu, a1.x, s); s = dp4a_ss((qw[1] >> 4) & 0x0f0f0f0fu, a1.y, s);
                s = dp4a_ss((qw[2] >> 4) & 0x0f0f0f0fu, a1.z, s); s = dp4a_ss((qw[3] >> 4) & 0x0f0f0f0fu, a1.w, s);
                const int bsum = (int)A.s[(size_t)c * A.s_stride + (size_t)it * 4 + bi];
                const int P = s - 8 * bsum;
                const float da = A.d[(size_t)c * A.d_stride + (size_t)it * 4 + bi];
                acc[c] += (dw * da) * (float)P;
                if (DBG && c == 0) B200_DBG_ADD(dbg.P + it * 4 + bi, P);
            }
        }
    }
}

template <int NC, bool DBG>
B200_HD void dot_item_q80(const uint8_t *rowp, int it, int nvalid, const ActView &A, float *acc, DbgSink dbg) {
    const uint8_t *p = rowp + (size_t)it * 136;
    uint32_t W[34];
    load_words<34>(p, W);    // block bi: half d at 34bi, int8 qs at 34bi+2
#pragma unroll
    for (int bi = 0; bi < 4; bi++) {
        if (bi < nvalid) {
            const int o = 34 * bi;
            const uint32_t dbits = (o & 2) ? (W[o >> 2] >> 16) : (W[o >> 2] & 0xffff);
            uint32_t qw[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int t = o + 2 + 4 * k;
                qw[k] = (t & 2) ? align2_word(W[t >> 2], W[(t >> 2) + 1 < 34 ? (t >> 2) + 1 : 33], 2) : W[t >> 2];
            }
            const float dw = h2f_bits(dbits);
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int8_t *ap = A.q + (size_t)c * A.q_stride + (size_t)it * act_item_stride<T_Q8_0>() + 32 * bi;
                const U4 a0 = ld128(ap), a1 = ld128(ap + 16);
                int s = 0;
                s = dp4a_ss(qw[0], a0.x, s); s = dp4a_ss(qw[1], a0.y, s); s = dp4a_ss(qw[2], a0.z, s); s = dp4a_ss(qw[3], a0.w, s);
                s = dp4a_ss(qw[4], a1.x, s); s = dp4a_ss(qw[5], a1.y, s); s = dp4a_ss(qw[6], a1.z, s); s = dp4a_ss(qw[7], a1.w, s);
                const float da = A.d[(size_t)c * A.d_stride + (size_t)it * 4 + bi];
                acc[c] += (dw * da) * (float)s;
                if (DBG && c == 0) B200_DBG_ADD(dbg.P + it * 4 + bi, s);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- dispatch
template <int TYPE> B200_HD int num_items(int K) { return (K + Traits<TYPE>::ITEM - 1) / Traits<TYPE>::ITEM; }

template <int TYPE, int NC, bool DBG>
B200_HD void dot_item(const uint8_t *rowp, int it, int K, const ActView &A, float *acc, DbgSink dbg) {
    if (TYPE == T_Q4_K || TYPE == T_Q5_K) dot_item_q45k<TYPE, NC, DBG>(rowp, it, A, acc, dbg);
    else if (TYPE == T_Q6_K) dot_item_q6k<NC, DBG>(rowp, it, A, acc, dbg);
    else {
        const int left = K / 32 - it * 4;
        const int nvalid = left < 4 ? left : 4;
        if (TYPE == T_Q4_0) dot_item_q40<NC, DBG>(rowp, it, nvalid, A, acc, dbg);
        else dot_item_q80<NC, DBG>(rowp, it, nvalid, A, acc, dbg);
    }
}

}  // namespace gemv
