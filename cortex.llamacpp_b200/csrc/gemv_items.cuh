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
// Split in two halves so the kernel can keep the activation registers of an item live while it walks
// several weight rows (bs1) or keep the decoded weights live while it walks several columns (bs2..4):
//     load_act<TYPE>(A, col, it)            -> ActRegs<TYPE>
//     item_dot<TYPE,DBG>(rowp, it, K, act)  -> float contribution of this item to (row, col)
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
B200_HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { uint32_t u; float f; } cv; cv.u = u; return cv.f;
#endif
}

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
template <int TYPE> B200_HD int num_items(int K) { return (K + Traits<TYPE>::ITEM - 1) / Traits<TYPE>::ITEM; }

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

// debug sink: exact integer pieces per weight block of the current row (block-sum test hook)
struct DbgSink {
    int32_t *P, *M;
};
#ifdef __CUDA_ARCH__
#define B200_DBG_ADD(ptr, v) atomicAdd((int *)(ptr), (int)(v))
#else
#define B200_DBG_ADD(ptr, v) (*(ptr) += (v))
#endif

// ---- activation registers of one item ---------------------------------------------------------------
template <int TYPE> struct ActRegs {        // 128-element items (Q6_K, Q4_0, Q8_0)
    U4 a[8];          // 128 int8
    U4 s;             // Q6_K: 8 bsums (int16) ; Q4_0: s.x,s.y = 4 block sums (int16) ; Q8_0: unused
    U4 d;             // Q6_K: d.x = block scale (f32 bits) ; Q4_0/Q8_0: 4 block scales (f32 bits)
};
template <> struct ActRegs<T_Q4_K> { U4 a[4]; uint32_t s01, s23; float da; };
template <> struct ActRegs<T_Q5_K> { U4 a[4]; uint32_t s01, s23; float da; };

template <int TYPE> B200_HD void load_act(const ActView &A, int c, int it, ActRegs<TYPE> &r) {
    const int8_t *ap = A.q + (size_t)c * A.q_stride + (size_t)it * act_item_stride<TYPE>();
    if constexpr (TYPE == T_Q4_K || TYPE == T_Q5_K) {
        r.a[0] = ld128(ap); r.a[1] = ld128(ap + 16); r.a[2] = ld128(ap + 32); r.a[3] = ld128(ap + 48);
        const int16_t *sp = A.s + (size_t)c * A.s_stride + (size_t)it * 4;
        r.s01 = ld32(sp); r.s23 = ld32(sp + 2);
        r.da = A.d[(size_t)c * A.d_stride + (it >> 2)];
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) r.a[i] = ld128(ap + 16 * i);
        if constexpr (TYPE == T_Q6_K) {
            r.s = ld128(A.s + (size_t)c * A.s_stride + (size_t)it * 8);
            r.d.x = ld32(A.d + (size_t)c * A.d_stride + (it >> 1));
        } else {
            if constexpr (TYPE == T_Q4_0) {
                const int16_t *sp = A.s + (size_t)c * A.s_stride + (size_t)it * 4;
                r.s.x = ld32(sp); r.s.y = ld32(sp + 2);
            }
            r.d = ld128(A.d + (size_t)c * A.d_stride + (size_t)it * 4);
        }
    }
}

// ---------------------------------------------------------------------------------------------- Q4_K / Q5_K
template <int TYPE, bool DBG>
B200_HD float item_dot_q45k(const uint8_t *rowp, int it, const ActRegs<TYPE> &r, DbgSink dbg) {
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
    int sA = 0, sB = 0, sA2 = 0, sB2 = 0;     // two chains each for ILP; integer adds are exact in any order
    sA  = dp4a_ss(lo[0], r.a[0].x, sA);  sA2 = dp4a_ss(lo[1], r.a[0].y, sA2); sA  = dp4a_ss(lo[2], r.a[0].z, sA);  sA2 = dp4a_ss(lo[3], r.a[0].w, sA2);
    sA  = dp4a_ss(lo[4], r.a[1].x, sA);  sA2 = dp4a_ss(lo[5], r.a[1].y, sA2); sA  = dp4a_ss(lo[6], r.a[1].z, sA);  sA2 = dp4a_ss(lo[7], r.a[1].w, sA2);
    sB  = dp4a_ss(hi[0], r.a[2].x, sB);  sB2 = dp4a_ss(hi[1], r.a[2].y, sB2); sB  = dp4a_ss(hi[2], r.a[2].z, sB);  sB2 = dp4a_ss(hi[3], r.a[2].w, sB2);
    sB  = dp4a_ss(hi[4], r.a[3].x, sB);  sB2 = dp4a_ss(hi[5], r.a[3].y, sB2); sB  = dp4a_ss(hi[6], r.a[3].z, sB);  sB2 = dp4a_ss(hi[7], r.a[3].w, sB2);
    const int P = sc0 * (sA + sA2) + sc1 * (sB + sB2);
    const int M = m0 * ((int)(int16_t)(r.s01 & 0xffff) + (int)(int16_t)(r.s01 >> 16)) +
                  m1 * ((int)(int16_t)(r.s23 & 0xffff) + (int)(int16_t)(r.s23 >> 16));
    if (DBG) { B200_DBG_ADD(dbg.P + (it >> 2), P); B200_DBG_ADD(dbg.M + (it >> 2), M); }
    return (d * r.da) * (float)P - (dmin * r.da) * (float)M;
}

// ---------------------------------------------------------------------------------------------- 2-byte aligned streams
B200_HD uint32_t half_at(const uint8_t *p) { return *(const uint16_t *)p; }

// loads N words starting at the 2-byte aligned p with N+1 aligned 32-bit loads
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
template <bool DBG>
B200_HD float item_dot_q6k(const uint8_t *rowp, int it, const ActRegs<T_Q6_K> &r, DbgSink dbg) {
    const uint8_t *b = rowp + (size_t)(it >> 1) * 210;
    const int h = it & 1;
    uint32_t L[16], H[8], S[2];
    load_words<16>(b + 64 * h, L);
    load_words<8>(b + 128 + 32 * h, H);
    load_words<2>(b + 192 + 8 * h, S);
    const float d = h2f_bits(half_at(b + 208));
    const uint32_t bsw[4] = {r.s.x, r.s.y, r.s.z, r.s.w};
    int P = 0;
    // scale group sg = 2t + (i>>2) covers word i (0..7) of quadrant t; 6-bit quants kept unsigned (0..63), the
    // -32 offset is applied through the activation bsums: sum (q-32)*a = sum q*a - 32*bsum
#pragma unroll
    for (int sg = 0; sg < 8; sg++) {
        const int t = sg >> 1;
        const U4 a = r.a[sg];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
        int s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = (sg & 1) * 4 + k;
            uint32_t q;
            if (t == 0)      q = (L[i] & 0x0f0f0f0fu)            | ((H[i] << 4) & 0x30303030u);
            else if (t == 1) q = (L[8 + i] & 0x0f0f0f0fu)        | ((H[i] << 2) & 0x30303030u);
            else if (t == 2) q = ((L[i] >> 4) & 0x0f0f0f0fu)     | (H[i] & 0x30303030u);
            else             q = ((L[8 + i] >> 4) & 0x0f0f0f0fu) | ((H[i] >> 2) & 0x30303030u);
            s = dp4a_ss(q, aw[k], s);
        }
        const int bsum = (int)(int16_t)((sg & 1) ? (bsw[sg >> 1] >> 16) : (bsw[sg >> 1] & 0xffff));
        const int scale = (int)(int8_t)((S[sg >> 2] >> (8 * (sg & 3))) & 0xff);
        P += scale * (s - 32 * bsum);
    }
    if (DBG) B200_DBG_ADD(dbg.P + (it >> 1), P);
    return (d * u2f(r.d.x)) * (float)P;
}

// ---------------------------------------------------------------------------------------------- Q4_0 / Q8_0
// item = up to 4 consecutive 32-element blocks; nvalid = number of blocks of this item inside K
template <bool DBG>
B200_HD float item_dot_q40(const uint8_t *rowp, int it, int nvalid, const ActRegs<T_Q4_0> &r, DbgSink dbg) {
    const uint8_t *p = rowp + (size_t)it * 72;
    uint32_t W[18];
    load_words<18>(p, W);    // stream of 72 bytes; block bi: half d at 18bi, qs at 18bi+2
    const uint32_t dav[4] = {r.d.x, r.d.y, r.d.z, r.d.w};
    float acc = 0.0f;
#pragma unroll
    for (int bi = 0; bi < 4; bi++) {
        if (bi < nvalid) {
            const int o = 18 * bi;
            const uint32_t dbits = (o & 2) ? (W[o >> 2] >> 16) : (W[o >> 2] & 0xffff);
            uint32_t qw[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int t = o + 2 + 4 * k;
                qw[k] = (t & 2) ? align2_word(W[t >> 2], W[(t >> 2) + 1 < 18 ? (t >> 2) + 1 : 17], 2) : W[t >> 2];
            }
            const U4 a0 = r.a[2 * bi], a1 = r.a[2 * bi + 1];
            int s = 0, s2 = 0;
            s  = dp4a_ss(qw[0] & 0x0f0f0f0fu, a0.x, s);  s2 = dp4a_ss(qw[1] & 0x0f0f0f0fu, a0.y, s2);
            s  = dp4a_ss(qw[2] & 0x0f0f0f0fu, a0.z, s);  s2 = dp4a_ss(qw[3] & 0x0f0f0f0fu, a0.w, s2);
            s  = dp4a_ss((qw[0] >> 4) & 0x0f0f0f0fu, a1.x, s);  s2 = dp4a_ss((qw[1] >> 4) & 0x0f0f0f0fu, a1.y, s2);
            s  = dp4a_ss((qw[2] >> 4) & 0x0f0f0f0fu, a1.z, s);  s2 = dp4a_ss((qw[3] >> 4) & 0x0f0f0f0fu, a1.w, s2);
            const uint32_t sw = bi < 2 ? r.s.x : r.s.y;
            const int bsum = (int)(int16_t)((bi & 1) ? (sw >> 16) : (sw & 0xffff));
            const int P = s + s2 - 8 * bsum;
            acc += (h2f_bits(dbits) * u2f(dav[bi])) * (float)P;
            if (DBG) B200_DBG_ADD(dbg.P + it * 4 + bi, P);
        }
    }
    return acc;
}

template <bool DBG>
B200_HD float item_dot_q80(const uint8_t *rowp, int it, int nvalid, const ActRegs<T_Q8_0> &r, DbgSink dbg) {
    const uint8_t *p = rowp + (size_t)it * 136;
    uint32_t W[34];
    load_words<34>(p, W);    // block bi: half d at 34bi, int8 qs at 34bi+2
    const uint32_t dav[4] = {r.d.x, r.d.y, r.d.z, r.d.w};
    float acc = 0.0f;
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
            const U4 a0 = r.a[2 * bi], a1 = r.a[2 * bi + 1];
            int s = 0, s2 = 0;
            s  = dp4a_ss(qw[0], a0.x, s); s2 = dp4a_ss(qw[1], a0.y, s2); s  = dp4a_ss(qw[2], a0.z, s); s2 = dp4a_ss(qw[3], a0.w, s2);
            s  = dp4a_ss(qw[4], a1.x, s); s2 = dp4a_ss(qw[5], a1.y, s2); s  = dp4a_ss(qw[6], a1.z, s); s2 = dp4a_ss(qw[7], a1.w, s2);
            acc += (h2f_bits(dbits) * u2f(dav[bi])) * (float)(s + s2);
            if (DBG) B200_DBG_ADD(dbg.P + it * 4 + bi, s + s2);
        }
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------- dispatch
template <int TYPE, bool DBG>
B200_HD float item_dot(const uint8_t *rowp, int it, int K, const ActRegs<TYPE> &r, DbgSink dbg) {
    if constexpr (TYPE == T_Q4_K || TYPE == T_Q5_K) return item_dot_q45k<TYPE, DBG>(rowp, it, r, dbg);
    else if constexpr (TYPE == T_Q6_K) return item_dot_q6k<DBG>(rowp, it, r, dbg);
    else {
        const int left = K / 32 - it * 4;
        const int nvalid = left < 4 ? left : 4;
        if constexpr (TYPE == T_Q4_0) return item_dot_q40<DBG>(rowp, it, nvalid, r, dbg);
        else return item_dot_q80<DBG>(rowp, it, nvalid, r, dbg);
    }
}

}  // namespace gemv
