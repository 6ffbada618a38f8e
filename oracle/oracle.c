/* oracle.c -- TEST INFRASTRUCTURE ONLY.  See oracle.h.
 *
 * A scalar, plain-C restatement of what the reference ggml CPU backend computes on the
 * hot path.  Written from the algorithm descriptions (block layouts in ggml-common.h:161-328),
 * not copied: loops are organised per logical sub-block, integer sums are kept exact and the
 * float combination is done in block order.
 *
 * Compile with -ffp-contract=off so that a*b+c is never fused (the reference's scalar code
 * is compiled the same way in its non-FMA positions; the float order difference to the AVX2
 * build is bounded in tests/test_oracle_pin.py).
 */
#define _GNU_SOURCE
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define QK   32
#define QKK  256

/* ---------------------------------------------------------------- fp16 */

float orc_f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp  = (h >> 10) & 0x1Fu;
    uint32_t man  = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {                       /* subnormal: renormalise */
            int e = -1;
            do { e++; man <<= 1; } while (!(man & 0x400u));
            man &= 0x3FFu;
            bits = sign | (uint32_t)(127 - 15 - e) << 23 | man << 13;
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | man << 13;
    } else {
        bits = sign | (exp + 112u) << 23 | man << 13;
    }
    float f; memcpy(&f, &bits, 4); return f;
}

uint16_t orc_f32_to_f16(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax   = x & 0x7FFFFFFFu;
    if (ax >= 0x7F800000u) {                     /* inf / nan */
        return (uint16_t)(sign | 0x7C00u | (ax > 0x7F800000u ? 0x200u | ((ax >> 13) & 0x3FFu) : 0));
    }
    if (ax >= 0x477FF000u) {                     /* rounds to >= 65520 -> inf */
        return (uint16_t)(sign | 0x7C00u);
    }
    if (ax < 0x38800000u) {                      /* result is subnormal or zero */
        if (ax < 0x33000000u) return (uint16_t)sign;         /* < 2^-25 -> 0 */
        int e = (int)(ax >> 23);                             /* biased exponent */
        uint32_t man = (ax & 0x7FFFFFu) | 0x800000u;
        int shift = 126 - e;                                 /* 14..24 */
        uint32_t q = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1u);
        uint32_t half = 1u << (shift - 1);
        if (rem > half || (rem == half && (q & 1u))) q++;
        return (uint16_t)(sign | q);
    }
    uint32_t man = ax & 0x7FFFFFu;
    uint32_t e   = (ax >> 23) - 112u;
    uint32_t q   = (e << 10) | (man >> 13);
    uint32_t rem = man & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) q++;   /* may carry into exponent: fine */
    return (uint16_t)(sign | q);
}

/* ---------------------------------------------------------------- layout helpers */

int orc_block_elems(int type) {
    switch (type) {
        case ORC_TYPE_Q4_0: case ORC_TYPE_Q8_0: return QK;
        case ORC_TYPE_Q4_K: case ORC_TYPE_Q5_K: case ORC_TYPE_Q6_K: case ORC_TYPE_Q8_K: return QKK;
        default: return 1;
    }
}

static size_t block_bytes(int type) {
    switch (type) {
        case ORC_TYPE_F32:  return 4;
        case ORC_TYPE_F16:  return 2;
        case ORC_TYPE_Q4_0: return 18;
        case ORC_TYPE_Q8_0: return 34;
        case ORC_TYPE_Q4_K: return 144;
        case ORC_TYPE_Q5_K: return 176;
        case ORC_TYPE_Q6_K: return 210;
        case ORC_TYPE_Q8_K: return 292;
        default: return 0;
    }
}

size_t orc_row_size(int type, int64_t k) {
    return (size_t)(k / orc_block_elems(type)) * block_bytes(type);
}

static inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline float    rdf32(const uint8_t *p) { float v; memcpy(&v, p, 4); return v; }
static inline int16_t  rds16(const uint8_t *p) { int16_t v; memcpy(&v, p, 2); return v; }

/* round-half-even of a float already scaled into int range.
 * (nearest_int, ggml-quants.c:372-377, is the same function via the 1.5*2^23 trick;
 *  _mm256_round_ps(_MM_ROUND_NEAREST) in the AVX2 q8_0 quantiser likewise.) */
static inline int rne(float v) { return (int)nearbyintf(v); }

/* ---------------------------------------------------------------- activation quantisers */

/* ggml-cpu-quants.c:808-860: d = amax/127 (stored fp16), id = 127/amax, q = RNE(x*id) */
void orc_quantize_row_q8_0(const float *x, void *vy, int64_t k) {
    uint8_t *y = (uint8_t *)vy;
    for (int64_t b = 0; b < k / QK; b++, y += 34, x += QK) {
        float amax = 0.0f;
        for (int j = 0; j < QK; j++) { float a = fabsf(x[j]); if (a > amax) amax = a; }
        const float d  = amax / 127.0f;
        const float id = amax != 0.0f ? 127.0f / amax : 0.0f;
        uint16_t dh = orc_f32_to_f16(d);
        memcpy(y, &dh, 2);
        for (int j = 0; j < QK; j++) y[2 + j] = (uint8_t)(int8_t)rne(x[j] * id);
    }
}

/* ggml-quants.c:2479-2513: iscale = -127/max (max = signed value of the max-abs element),
 * q = min(127, RNE(iscale*x)), d = 1/iscale (f32), bsums per 16 */
void orc_quantize_row_q8_K(const float *x, void *vy, int64_t k) {
    uint8_t *y = (uint8_t *)vy;
    for (int64_t b = 0; b < k / QKK; b++, y += 292, x += QKK) {
        float max = 0.0f, amax = 0.0f;
        for (int j = 0; j < QKK; j++) { float a = fabsf(x[j]); if (a > amax) { amax = a; max = x[j]; } }
        if (amax == 0.0f) { memset(y, 0, 292); continue; }   /* reference leaves bsums untouched; zeros are what a fresh buffer holds */
        const float iscale = -127.0f / max;
        int8_t *qs = (int8_t *)(y + 4);
        for (int j = 0; j < QKK; j++) { int v = rne(iscale * x[j]); qs[j] = (int8_t)(v > 127 ? 127 : v); }
        for (int j = 0; j < 16; j++) {
            int s = 0;
            for (int i = 0; i < 16; i++) s += qs[16 * j + i];
            int16_t s16 = (int16_t)s; memcpy(y + 260 + 2 * j, &s16, 2);
        }
        const float d = 1.0f / iscale;
        memcpy(y, &d, 4);
    }
}

/* ggml-quants.c:35-70 (quantize_row_q4_0_ref): d = max/-8 where max is the signed max-abs
 * element, q = min(15, (int8)(x*id + 8.5)) */
void orc_quantize_row_q4_0(const float *x, void *vy, int64_t k) {
    uint8_t *y = (uint8_t *)vy;
    for (int64_t b = 0; b < k / QK; b++, y += 18, x += QK) {
        float amax = 0.0f, max = 0.0f;
        for (int j = 0; j < QK; j++) { float a = fabsf(x[j]); if (a > amax) { amax = a; max = x[j]; } }
        const float d  = max / -8.0f;
        const float id = d != 0.0f ? 1.0f / d : 0.0f;
        uint16_t dh = orc_f32_to_f16(d);
        memcpy(y, &dh, 2);
        for (int j = 0; j < QK / 2; j++) {
            float x0 = x[j] * id, x1 = x[QK / 2 + j] * id;
            int a = (int8_t)(x0 + 8.5f), c = (int8_t)(x1 + 8.5f);
            if (a > 15) a = 15; if (c > 15) c = 15;
            y[2 + j] = (uint8_t)(a | (c << 4));
        }
    }
}

/* ---------------------------------------------------------------- K-quant scale unpack */

/* get_scale_min_k4, ggml-quants.c:631-639: 8 six-bit (scale,min) pairs packed in 12 bytes */
static void unpack_scales_k4(const uint8_t *s, int sc[8], int mn[8]) {
    for (int j = 0; j < 4; j++) {
        sc[j]     = s[j] & 63;
        mn[j]     = s[j + 4] & 63;
        sc[j + 4] = (s[j + 8] & 0x0F) | ((s[j] >> 6) << 4);
        mn[j + 4] = (s[j + 8] >> 4)   | ((s[j + 4] >> 6) << 4);
    }
}

/* unsigned quant value of element e (0..255) in a q4_K / q5_K block */
static inline int q4k_elem(const uint8_t *qs, int e) {
    int g = e >> 6, r = e & 63;            /* 64-element group shares 32 bytes */
    uint8_t byte = qs[32 * g + (r & 31)];
    return r < 32 ? (byte & 0x0F) : (byte >> 4);
}
static inline int q5k_elem(const uint8_t *qh, const uint8_t *qs, int e) {
    int j = e >> 5;                        /* sub-block index = bit index inside qh byte */
    return q4k_elem(qs, e) | (((qh[e & 31] >> j) & 1) << 4);
}
/* signed quant value (q-32) of element e in a q6_K block: ggml-quants.c:1690-1719 */
static inline int q6k_elem(const uint8_t *ql, const uint8_t *qh, int e) {
    int h = e >> 7, r = e & 127, t = r >> 5, l = r & 31;
    uint8_t lb = ql[64 * h + l + 32 * (t & 1)];
    int lo = t < 2 ? (lb & 0x0F) : (lb >> 4);
    int hi = (qh[32 * h + l] >> (2 * t)) & 3;
    return (lo | (hi << 4)) - 32;
}

/* ---------------------------------------------------------------- dequantise */

void orc_dequantize_row(int type, const void *vx, float *y, int64_t k) {
    const uint8_t *x = (const uint8_t *)vx;
    switch (type) {
    case ORC_TYPE_F32: memcpy(y, x, (size_t)k * 4); break;
    case ORC_TYPE_F16: for (int64_t i = 0; i < k; i++) y[i] = orc_f16_to_f32(rd16(x + 2 * i)); break;
    case ORC_TYPE_Q4_0:
        for (int64_t b = 0; b < k / QK; b++, x += 18, y += QK) {
            const float d = orc_f16_to_f32(rd16(x));
            for (int j = 0; j < 16; j++) {
                y[j]      = (float)((x[2 + j] & 0x0F) - 8) * d;
                y[j + 16] = (float)((x[2 + j] >> 4) - 8) * d;
            }
        }
        break;
    case ORC_TYPE_Q8_0:
        for (int64_t b = 0; b < k / QK; b++, x += 34, y += QK) {
            const float d = orc_f16_to_f32(rd16(x));
            for (int j = 0; j < QK; j++) y[j] = (float)(int8_t)x[2 + j] * d;
        }
        break;
    case ORC_TYPE_Q4_K:
        for (int64_t b = 0; b < k / QKK; b++, x += 144, y += QKK) {
            const float d = orc_f16_to_f32(rd16(x)), dmin = orc_f16_to_f32(rd16(x + 2));
            int sc[8], mn[8]; unpack_scales_k4(x + 4, sc, mn);
            for (int e = 0; e < QKK; e++) {
                const float d1 = d * sc[e >> 5], m1 = dmin * mn[e >> 5];
                y[e] = d1 * q4k_elem(x + 16, e) - m1;
            }
        }
        break;
    case ORC_TYPE_Q5_K:
        for (int64_t b = 0; b < k / QKK; b++, x += 176, y += QKK) {
            const float d = orc_f16_to_f32(rd16(x)), dmin = orc_f16_to_f32(rd16(x + 2));
            int sc[8], mn[8]; unpack_scales_k4(x + 4, sc, mn);
            for (int e = 0; e < QKK; e++) {
                const float d1 = d * sc[e >> 5], m1 = dmin * mn[e >> 5];
                y[e] = d1 * q5k_elem(x + 16, x + 48, e) - m1;
            }
        }
        break;
    case ORC_TYPE_Q6_K:
        for (int64_t b = 0; b < k / QKK; b++, x += 210, y += QKK) {
            const float d = orc_f16_to_f32(rd16(x + 208));
            const int8_t *scales = (const int8_t *)(x + 192);
            for (int e = 0; e < QKK; e++) {
                int h = e >> 7, t = (e & 127) >> 5, l = e & 31;
                y[e] = d * scales[8 * h + 2 * t + l / 16] * q6k_elem(x, x + 128, e);
            }
        }
        break;
    default: break;
    }
}

/* ---------------------------------------------------------------- exact block sums */

void orc_block_sums(int type, const void *vw, const void *vact, int64_t k, int32_t *P, int32_t *M) {
    const uint8_t *w = (const uint8_t *)vw, *a = (const uint8_t *)vact;
    switch (type) {
    case ORC_TYPE_Q4_0:
        for (int64_t b = 0; b < k / QK; b++, w += 18, a += 34) {
            const int8_t *q8 = (const int8_t *)(a + 2);
            int s = 0;
            for (int j = 0; j < 16; j++) s += ((w[2 + j] & 0x0F) - 8) * q8[j] + ((w[2 + j] >> 4) - 8) * q8[j + 16];
            P[b] = s; if (M) M[b] = 0;
        }
        break;
    case ORC_TYPE_Q8_0:
        for (int64_t b = 0; b < k / QK; b++, w += 34, a += 34) {
            const int8_t *q8 = (const int8_t *)(a + 2), *qw = (const int8_t *)(w + 2);
            int s = 0;
            for (int j = 0; j < QK; j++) s += qw[j] * q8[j];
            P[b] = s; if (M) M[b] = 0;
        }
        break;
    case ORC_TYPE_Q4_K: case ORC_TYPE_Q5_K: {
        const size_t wb = type == ORC_TYPE_Q4_K ? 144 : 176;
        for (int64_t b = 0; b < k / QKK; b++, w += wb, a += 292) {
            const int8_t *q8 = (const int8_t *)(a + 4);
            int sc[8], mn[8]; unpack_scales_k4(w + 4, sc, mn);
            int p = 0, m = 0;
            for (int j = 0; j < 8; j++) {
                int s = 0;
                for (int l = 0; l < 32; l++) {
                    int e = 32 * j + l;
                    int q = type == ORC_TYPE_Q4_K ? q4k_elem(w + 16, e) : q5k_elem(w + 16, w + 48, e);
                    s += q * q8[e];
                }
                p += sc[j] * s;
            }
            for (int j = 0; j < 16; j++) m += rds16(a + 260 + 2 * j) * mn[j / 2];
            P[b] = p; if (M) M[b] = m;
        }
        break; }
    case ORC_TYPE_Q6_K:
        for (int64_t b = 0; b < k / QKK; b++, w += 210, a += 292) {
            const int8_t *q8 = (const int8_t *)(a + 4);
            const int8_t *scales = (const int8_t *)(w + 192);
            int p = 0;
            for (int e = 0; e < QKK; e++) {
                int h = e >> 7, t = (e & 127) >> 5, l = e & 31;
                p += scales[8 * h + 2 * t + l / 16] * q6k_elem(w, w + 128, e) * q8[e];
            }
            P[b] = p; if (M) M[b] = 0;
        }
        break;
    default: break;
    }
}

/* ---------------------------------------------------------------- the reference's SIMD summation order
 * The AVX2 branches of ggml_vec_dot_{q4_0,q8_0}_q8_0 (ggml-cpu-quants.c:2273-2296, 3935-3952) and
 * ggml_vec_dot_{q4_K,q5_K,q6_K}_q8_K (:6776-6837, 7413-7490, 8405-8481) all have the same shape: per weight block an
 * 8-lane int32 vector whose lane l collects the products of bytes 4l..4l+3 of every 32-element group of the block
 * (maddubs pairs, then madd with the group's sub-scale), converted to float and accumulated with one FMA per block into an
 * 8-lane f32 accumulator (acc = fma(d_block, (float)lanes, acc)), reduced at the end by hsum_float_8 (:49-55):
 * ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)).  The mins of q4_K go through a 4-lane FMA accumulator (:6796-6799, 6833-6836), the
 * mins of q5_K through a scalar float (:7439-7443).  Restating exactly that order makes the float result bit-identical to
 * the reference build in oracle/_ref (which is what the tests assert), not just equal up to summation order. */
static void lane_sums(int type, const uint8_t *w, const uint8_t *a, int32_t L[8], int32_t prod[4]) {
    for (int l = 0; l < 8; l++) L[l] = 0;
    for (int k = 0; k < 4; k++) prod[k] = 0;
    switch (type) {
    case ORC_TYPE_Q4_0: {
        const int8_t *q8 = (const int8_t *)(a + 2);
        for (int i = 0; i < 32; i++) {
            const int q = i < 16 ? (w[2 + i] & 0x0F) - 8 : (w[2 + i - 16] >> 4) - 8;
            L[i >> 2] += q * q8[i];
        }
        break; }
    case ORC_TYPE_Q8_0: {
        const int8_t *q8 = (const int8_t *)(a + 2), *qw = (const int8_t *)(w + 2);
        for (int i = 0; i < 32; i++) L[i >> 2] += qw[i] * q8[i];
        break; }
    case ORC_TYPE_Q4_K: case ORC_TYPE_Q5_K: {
        const int8_t *q8 = (const int8_t *)(a + 4);
        int sc[8], mn[8]; unpack_scales_k4(w + 4, sc, mn);
        for (int j = 0; j < 8; j++)
            for (int i = 0; i < 32; i++) {
                const int e = 32 * j + i;
                const int q = type == ORC_TYPE_Q4_K ? q4k_elem(w + 16, e) : q5k_elem(w + 16, w + 48, e);
                L[i >> 2] += sc[j] * (q * q8[e]);
            }
        for (int k = 0; k < 4; k++) {       /* q8s = hadd_epi16(bsums lo, hi) = per-32 sums; prod = madd_epi16(mins, q8s) */
            const int s0 = (int16_t)(rds16(a + 260 + 2 * (4 * k)) + rds16(a + 260 + 2 * (4 * k + 1)));
            const int s1 = (int16_t)(rds16(a + 260 + 2 * (4 * k + 2)) + rds16(a + 260 + 2 * (4 * k + 3)));
            prod[k] = mn[2 * k] * s0 + mn[2 * k + 1] * s1;
        }
        break; }
    case ORC_TYPE_Q6_K: {
        const int8_t *q8 = (const int8_t *)(a + 4);
        const int8_t *scales = (const int8_t *)(w + 192);
        for (int e = 0; e < QKK; e++) {
            const int h = e >> 7, t = (e & 127) >> 5, i = e & 31;
            L[i >> 2] += scales[8 * h + 2 * t + i / 16] * (q6k_elem(w, w + 128, e) * q8[e]);
        }
        break; }
    default: break;
    }
}

float orc_vec_dot(int type, const void *vw, const void *vact, int64_t k) {
    const int be = orc_block_elems(type);
    const int64_t nb = k / be;
    const uint8_t *w = (const uint8_t *)vw, *a = (const uint8_t *)vact;
    const size_t wb = type == ORC_TYPE_Q4_0 ? 18 : type == ORC_TYPE_Q8_0 ? 34 : type == ORC_TYPE_Q4_K ? 144 : type == ORC_TYPE_Q5_K ? 176 : 210;
    const size_t ab = be == 32 ? 34 : 292;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, acc_m[4] = {0, 0, 0, 0}, summs = 0.0f;
    for (int64_t b = 0; b < nb; b++, w += wb, a += ab) {
        int32_t L[8], prod[4];
        lane_sums(type, w, a, L, prod);
        float d;
        if (be == 32) d = orc_f16_to_f32(rd16(w)) * orc_f16_to_f32(rd16(a));
        else d = rdf32(a) * orc_f16_to_f32(rd16(w + (type == ORC_TYPE_Q6_K ? 208 : 0)));
        for (int l = 0; l < 8; l++) acc[l] = fmaf(d, (float)L[l], acc[l]);
        if (type == ORC_TYPE_Q4_K || type == ORC_TYPE_Q5_K) {
            const float dmin = -rdf32(a) * orc_f16_to_f32(rd16(w + 2));
            if (type == ORC_TYPE_Q4_K) for (int kk = 0; kk < 4; kk++) acc_m[kk] = fmaf(dmin, (float)prod[kk], acc_m[kk]);
            else summs += dmin * (float)((prod[0] + prod[1]) + (prod[2] + prod[3]));      /* hadd_epi32 twice, then scalar float (:7442-7443) */
        }
    }
    const float h8 = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
    if (type == ORC_TYPE_Q4_K) return h8 + ((acc_m[0] + acc_m[2]) + (acc_m[1] + acc_m[3]));
    if (type == ORC_TYPE_Q5_K) return h8 + summs;
    return h8;
}

/* vec_dot_type table: ggml-cpu.c:266-341 */
static int act_type_for(int type) {
    return (type == ORC_TYPE_Q4_0 || type == ORC_TYPE_Q8_0) ? ORC_TYPE_Q8_0 : ORC_TYPE_Q8_K;
}

/* ggml_vec_dot_f32 as the AVX2 + FMA build computes it (ggml-cpu.c:1454-1494 with GGML_F32_STEP 32, GGML_F32_EPR 8): four 8-lane
 * accumulators, one FMA per element, GGML_F32x8_REDUCE (:671-689), then the n % 32 leftovers as float mul + add */
static float dot_f32(const float *a, const float *b, int64_t n) {
    float sum[4][8];
    for (int j = 0; j < 4; j++) for (int l = 0; l < 8; l++) sum[j][l] = 0.0f;
    const int64_t np = n & ~(int64_t)31;
    for (int64_t i = 0; i < np; i += 32)
        for (int j = 0; j < 4; j++)
            for (int l = 0; l < 8; l++) sum[j][l] = fmaf(a[i + 8 * j + l], b[i + 8 * j + l], sum[j][l]);
    float x0[8], t[4];
    for (int l = 0; l < 8; l++) x0[l] = (sum[0][l] + sum[2][l]) + (sum[1][l] + sum[3][l]);
    for (int l = 0; l < 4; l++) t[l] = x0[l] + x0[l + 4];
    float sumf = (t[0] + t[1]) + (t[2] + t[3]);
    for (int64_t i = np; i < n; i++) { const float pr = a[i] * b[i]; sumf += pr; }
    return sumf;
}

void orc_mul_mat(int type, const void *W, const float *x, float *dst, int64_t N, int64_t K, int64_t Mcols) {
    if (type == ORC_TYPE_F32 || type == ORC_TYPE_F16) {
        float *row = (float *)malloc(sizeof(float) * (size_t)K);
        for (int64_t n = 0; n < N; n++) {
            orc_dequantize_row(type, (const uint8_t *)W + (size_t)n * orc_row_size(type, K), row, K);
            for (int64_t m = 0; m < Mcols; m++) {
                if (type == ORC_TYPE_F16) {            /* CPU converts src1 to f16 first (vec_dot_type F16) */
                    double s = 0;
                    for (int64_t i = 0; i < K; i++) s += (double)row[i] * orc_f16_to_f32(orc_f32_to_f16(x[m * K + i]));
                    dst[m * N + n] = (float)s;
                } else dst[m * N + n] = dot_f32(row, x + m * K, K);
            }
        }
        free(row);
        return;
    }
    const int at = act_type_for(type);
    const size_t ars = orc_row_size(at, K), wrs = orc_row_size(type, K);
    uint8_t *act = (uint8_t *)calloc((size_t)Mcols, ars);
    for (int64_t m = 0; m < Mcols; m++) {
        if (at == ORC_TYPE_Q8_0) orc_quantize_row_q8_0(x + m * K, act + m * ars, K);
        else                     orc_quantize_row_q8_K(x + m * K, act + m * ars, K);
    }
    for (int64_t m = 0; m < Mcols; m++)
        for (int64_t n = 0; n < N; n++)
            dst[m * N + n] = orc_vec_dot(type, (const uint8_t *)W + (size_t)n * wrs, act + m * ars, K);
    free(act);
}

void orc_mul_mat_id(int type, const void *as, const float *b, const int32_t *ids, float *dst,
                    int64_t N, int64_t K, int64_t n_expert, int64_t n_used, int64_t n_tok, int64_t b_ne1) {
    (void)n_expert;
    const size_t ebytes = (size_t)N * orc_row_size(type, K);
    for (int64_t t = 0; t < n_tok; t++)
        for (int64_t u = 0; u < n_used; u++) {
            const int32_t e = ids[t * n_used + u];
            const float *xb = b + (t * b_ne1 + (u % b_ne1)) * K;
            orc_mul_mat(type, (const uint8_t *)as + (size_t)e * ebytes, xb, dst + (t * n_used + u) * N, N, K, 1);
        }
}

/* ---------------------------------------------------------------- glue ops */

/* ggml_compute_forward_rms_norm_f32: sum of squares accumulated in double (ggml_float), mean,
 * scale = 1/sqrtf(mean + eps) */
void orc_rms_norm(const float *x, float *y, int64_t ncols, int64_t nrows, float eps) {
    for (int64_t r = 0; r < nrows; r++, x += ncols, y += ncols) {
        double sum = 0.0;
        for (int64_t i = 0; i < ncols; i++) sum += (double)(x[i] * x[i]);
        const float mean = (float)(sum / (double)ncols);
        const float scale = 1.0f / sqrtf(mean + eps);
        for (int64_t i = 0; i < ncols; i++) y[i] = x[i] * scale;
    }
}

/* ggml.c:3735-3750 */
static float yarn_corr_dim(int n_dims, int n_ctx_orig, float n_rot, float base) {
    return n_dims * logf(n_ctx_orig / (n_rot * 2 * (float)M_PI)) / (2 * logf(base));
}

/* ggml-cpu.c:10573-10800: theta advances multiplicatively (theta *= theta_scale), YaRN mix */
void orc_rope(const float *x, float *y, const int32_t *pos, const float *ff,
              int64_t ne0, int64_t n_head, int64_t n_tok, int n_dims, int mode, int n_ctx_orig,
              float freq_base, float freq_scale, float ext_factor, float attn_factor,
              float beta_fast, float beta_slow) {
    const float theta_scale = powf(freq_base, -2.0f / n_dims);
    float lo = floorf(yarn_corr_dim(n_dims, n_ctx_orig, beta_fast, freq_base));
    float hi = ceilf(yarn_corr_dim(n_dims, n_ctx_orig, beta_slow, freq_base));
    if (lo < 0) lo = 0;
    if (hi > n_dims - 1) hi = (float)(n_dims - 1);
    const int neox = mode & 2;
    float *cs = (float *)malloc(sizeof(float) * (size_t)ne0);
    for (int64_t t = 0; t < n_tok; t++) {
        float theta = (float)pos[t];
        for (int64_t i0 = 0; i0 < ne0; i0 += 2) {
            const float f = ff ? ff[i0 / 2] : 1.0f;
            const float te = theta / f;
            float ti = freq_scale * te, th = ti, ms = attn_factor;
            if (ext_factor != 0.0f) {
                float yv = (i0 / 2 - lo) / fmaxf(0.001f, hi - lo);
                float ramp = (1 - fminf(1, fmaxf(0, yv))) * ext_factor;
                th = ti * (1 - ramp) + te * ramp;
                ms *= 1.0f + 0.1f * logf(1.0f / freq_scale);
            }
            cs[i0] = cosf(th) * ms; cs[i0 + 1] = sinf(th) * ms;
            theta *= theta_scale;
        }
        for (int64_t h = 0; h < n_head; h++) {
            const float *src = x + (t * n_head + h) * ne0;
            float *dst = y + (t * n_head + h) * ne0;
            for (int64_t i0 = 0; i0 < n_dims; i0 += 2) {
                const float c = cs[i0], s = cs[i0 + 1];
                if (!neox) {
                    const float x0 = src[i0], x1 = src[i0 + 1];
                    dst[i0] = x0 * c - x1 * s; dst[i0 + 1] = x0 * s + x1 * c;
                } else {
                    const int64_t ic = i0 / 2;
                    const float x0 = src[ic], x1 = src[ic + n_dims / 2];
                    dst[ic] = x0 * c - x1 * s; dst[ic + n_dims / 2] = x0 * s + x1 * c;
                }
            }
            for (int64_t i0 = n_dims; i0 < ne0; i0++) dst[i0] = src[i0];
        }
    }
    free(cs);
}

/* ggml_compute_forward_soft_max_f32 (ggml-cpu.c:10224-10320): y = softmax(x*scale + mask), max-subtracted.  The exponentials and
 * their sum follow ggml_vec_soft_max_f32's AVX2 path (:2261-2271, 2296-2300): full groups of 8 go through ggml_v_expf (polynomial,
 * not libm) and enter the double sum as ONE float, the horizontal sum ((v0+v4)+(v2+v6))+((v1+v5)+(v3+v7)); the n % 8 leftovers go
 * through expf one by one */
static float v_expf_lane(float x);
void orc_soft_max(const float *x, const uint16_t *mask, float *y, int64_t ncols, int64_t nrows, float scale) {
    for (int64_t r = 0; r < nrows; r++, x += ncols, y += ncols) {
        float mx = -INFINITY;
        for (int64_t i = 0; i < ncols; i++) {
            float w = x[i] * scale;
            if (mask) { const float mv = 1.0f * orc_f16_to_f32(mask[r * ncols + i]); w += mv; }
            y[i] = w;
            if (y[i] > mx) mx = y[i];
        }
        double sum = 0.0;
        int64_t i = 0;
        for (; i + 7 < ncols; i += 8) {
            float v[8];
            for (int l = 0; l < 8; l++) { v[l] = v_expf_lane(y[i + l] - mx); y[i + l] = v[l]; }
            const float s8 = ((v[0] + v[4]) + (v[2] + v[6])) + ((v[1] + v[5]) + (v[3] + v[7]));
            sum += (double)s8;
        }
        for (; i < ncols; i++) { y[i] = expf(y[i] - mx); sum += (double)y[i]; }
        const float inv = (float)(1.0 / sum);
        for (int64_t j = 0; j < ncols; j++) y[j] *= inv;
    }
}

/* silu as the SIMD builds of the CPU backend compute it: ggml_vec_silu_f32 (ggml-cpu.c:2221-2243) maps every full vector
 * through ggml_v_silu = x / (1 + ggml_v_expf(0 - x)) (:2116-2164), a polynomial expf (NOT libm); only the n % 8 leftovers go
 * through ggml_silu_f32 = x/(1+expf(-x)).  ggml_v_expf restated per lane (lanes are independent): */
static float v_expf_lane(float x) {
    union { float f; uint32_t u; } cz, ck, cs1, cs2;
    const float r = 0x1.8p23f;
    const float z = fmaf(x, 0x1.715476p+0f, r);
    const float n = z - r;
    const float b = fmaf(-n, 0x1.7f7d1cp-20f, fmaf(-n, 0x1.62e4p-1f, x));
    cz.f = z;
    const uint32_t e = cz.u << 23;
    ck.u = e + 0x3f800000u;
    const float k = ck.f, an = fabsf(n), u = b * b;
    const float j = fmaf(fmaf(fmaf(0x1.0e4020p-7f, b, 0x1.573e2ep-5f), u, fmaf(0x1.555e66p-3f, b, 0x1.fffdb6p-2f)), u, 0x1.ffffecp-1f * b);
    if (!(an > 126.0f)) return fmaf(j, k, k);
    const uint32_t g = n <= 0.0f ? 0x82000000u : 0u;
    cs1.u = g + 0x7f000000u; cs2.u = e - g;
    if (an > 192.0f) return cs1.f * cs1.f;
    return fmaf(cs2.f, j, cs2.f) * cs1.f;
}
void orc_silu_mul(const float *gate, const float *up, float *y, int64_t n) {
    const int64_t nv = n & ~(int64_t)7;
    for (int64_t i = 0; i < nv; i++) y[i] = (gate[i] / (1.0f + v_expf_lane(0.0f - gate[i]))) * up[i];
    for (int64_t i = nv; i < n; i++) y[i] = (gate[i] / (1.0f + expf(-gate[i]))) * up[i];
}

/* ---------------------------------------------------------------- flash attention */

/* ggml_vec_dot_f16 as the AVX2 + F16C + FMA build computes it (ggml-cpu.c:1565-1605 with GGML_F16_STEP 32, GGML_F16_EPR 8,
 * :704-705): four 8-lane f32 accumulators, one FMA per element, then GGML_F32x8_REDUCE (:671-689):
 * (s0+s2)+(s1+s3) lane-wise, low half + high half, two horizontal adds.  k: f16 bits, q: f32 values already rounded to f16 */
static float dot_f16_simd_order(const uint8_t *k, const float *q, int64_t n) {
    float sum[4][8];
    for (int j = 0; j < 4; j++) for (int l = 0; l < 8; l++) sum[j][l] = 0.0f;
    const int64_t np = n & ~(int64_t)31;
    for (int64_t i = 0; i < np; i += 32)
        for (int j = 0; j < 4; j++)
            for (int l = 0; l < 8; l++) {
                const int64_t e = i + 8 * j + l;
                sum[j][l] = fmaf(orc_f16_to_f32(rd16(k + 2 * e)), q[e], sum[j][l]);
            }
    float x0[8], t[4];
    for (int l = 0; l < 8; l++) x0[l] = (sum[0][l] + sum[2][l]) + (sum[1][l] + sum[3][l]);
    for (int l = 0; l < 4; l++) t[l] = x0[l] + x0[l + 4];
    double sumf = (double)((t[0] + t[1]) + (t[2] + t[3]));
    for (int64_t e = np; e < n; e++) sumf += (double)(orc_f16_to_f32(rd16(k + 2 * e)) * q[e]);
    return (float)sumf;
}

void orc_flash_attn_ext(const float *q, const void *k, const void *v, const uint16_t *mask, float *dst,
                        int64_t D, int64_t n_q, int64_t H, int64_t n_kv, int64_t Hkv,
                        int type_k, int type_v,
                        size_t k_nb1, size_t k_nb2, size_t v_nb1, size_t v_nb2, size_t mask_nb1,
                        float scale, float logit_softcap) {
    if (logit_softcap != 0.0f) scale /= logit_softcap;
    const int64_t gq = H / Hkv;
    float *acc = (float *)malloc(sizeof(float) * (size_t)D * 3);
    float *vrow = acc + D, *qf = acc + 2 * D;
    uint16_t *acc16 = (uint16_t *)malloc(2 * (size_t)D);
    uint8_t *qq = (uint8_t *)malloc(orc_row_size(ORC_TYPE_Q8_0, D) + 16);
    for (int64_t h = 0; h < H; h++)
        for (int64_t iq = 0; iq < n_q; iq++) {
            const float *pq = q + (h * n_q + iq) * D;
            /* Q -> K's vec_dot_type: f16 for f16 K, q8_0 for q8_0 / q4_0 K */
            if (type_k == ORC_TYPE_F16) for (int64_t d = 0; d < D; d++) qf[d] = orc_f16_to_f32(orc_f32_to_f16(pq[d]));
            else orc_quantize_row_q8_0(pq, qq, D);
            float S = 0.0f, Mx = -INFINITY;
            for (int64_t d = 0; d < D; d++) { acc[d] = 0.0f; acc16[d] = 0; }
            const uint16_t *mp = mask ? (const uint16_t *)((const uint8_t *)mask + iq * mask_nb1) : NULL;
            for (int64_t ic = 0; ic < n_kv; ic++) {
                const float mv = mp ? orc_f16_to_f32(mp[ic]) : 0.0f;
                if (mv == -INFINITY) continue;
                const uint8_t *kr = (const uint8_t *)k + ic * k_nb1 + (h / gq) * k_nb2;
                const uint8_t *vr = (const uint8_t *)v + ic * v_nb1 + (h / gq) * v_nb2;
                float s;
                if (type_k == ORC_TYPE_F16) s = dot_f16_simd_order(kr, qf, D);
                else s = orc_vec_dot(type_k, kr, qq, D);
                s *= scale;
                if (logit_softcap != 0.0f) s = logit_softcap * tanhf(s);
                s += mv;
                const float Mold = Mx;
                float ms = 1.0f, vs = 1.0f;
                if (s > Mx) { Mx = s; ms = expf(Mold - Mx); } else vs = expf(s - Mx);
                if (type_v == ORC_TYPE_F16) {
                    /* fp16 accumulator: scale then mad, each rounded to fp16 (ggml_vec_scale_f16 / ggml_vec_mad_f16) */
                    for (int64_t d = 0; d < D; d++) {
                        float a16 = orc_f16_to_f32(acc16[d]);
                        if (ms != 1.0f) a16 = orc_f16_to_f32(orc_f32_to_f16(a16 * ms));
                        a16 = fmaf(orc_f16_to_f32(rd16(vr + 2 * d)), vs, a16);      /* GGML_F16_VEC_FMA = _mm256_fmadd_ps (ggml-cpu.c:665, 1706) */
                        acc16[d] = orc_f32_to_f16(a16);
                    }
                } else {
                    orc_dequantize_row(type_v, vr, vrow, D);
                    /* ggml_vec_scale_f32 (mul) then ggml_vec_mad_f32 = GGML_F32_VEC_FMA (_mm256_fmadd_ps) */
                    for (int64_t d = 0; d < D; d++) { if (ms != 1.0f) acc[d] *= ms; acc[d] = fmaf(vrow[d], vs, acc[d]); }
                }
                S = S * ms + vs;
            }
            if (type_v == ORC_TYPE_F16) for (int64_t d = 0; d < D; d++) acc[d] = orc_f16_to_f32(acc16[d]);
            const float Sinv = 1.0f / S;
            float *o = dst + (iq * H + h) * D;
            for (int64_t d = 0; d < D; d++) o[d] = acc[d] * Sinv;
        }
    free(acc); free(acc16); free(qq);
}
