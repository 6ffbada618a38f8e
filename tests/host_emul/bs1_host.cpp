// Host emulation of the batch-1 GEMV block decoders (cortex.llamacpp_b200/csrc/gemv_bs1_items.cuh): the very same decode code
// the CUDA kernel runs per lane, compiled for the CPU, fed with the shared-memory activation layouts the kernel's prologue
// builds (272-byte padded int8 blocks + per-32 sums for Q4_K/Q5_K, 144-byte padded 128-weight items + per-16 sums for Q6_K),
// so the bit-twiddling (packed 6-bit scale unpack, in-place high nibbles, dp2a mins, PRMT re-alignment) can be checked against
// the oracle's exact integer block sums without a GPU.  TEST INFRASTRUCTURE.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../cortex.llamacpp_b200/csrc/gemv_bs1_items.cuh"

using namespace bs1;

// act: q8_K activations of one column: int8 q[K], float d[K/256], int16 bsums[K/16] (the reference's block_q8_K fields, split)
extern "C" int emul_bs1(int type, const uint8_t *W, size_t rb, int N, int K, const int8_t *q, const float *d, const int16_t *bsums,
                        float *dst, int32_t *P, int32_t *M, int phase) {
    const int nblk = K / 256;
    std::vector<uint8_t> aq64((size_t)nblk * 272 + 64, 0), aq128((size_t)(K / 128) * 144 + 64, 0);
    for (int e = 0; e < K; e++) {
        aq64[(size_t)(e >> 8) * 272 + (e & 255)] = (uint8_t)q[e];
        aq128[(size_t)(e >> 7) * 144 + (e & 127)] = (uint8_t)q[e];
    }
    std::vector<int16_t> s32((size_t)K / 32 + 16, 0), s16((size_t)K / 16 + 16, 0);
    for (int i = 0; i < K / 16; i++) s16[i] = bsums[i];
    for (int i = 0; i < K / 32; i++) s32[i] = (int16_t)(bsums[2 * i] + bsums[2 * i + 1]);
    std::vector<float> ad(nblk + 4, 0.0f);
    memcpy(ad.data(), d, (size_t)nblk * 4);
    // emulate a ring stage: the row sits at an address with the requested 2-byte phase (Q6_K rows are only 2-byte aligned)
    std::vector<uint8_t> stage(rb + 96);
    for (int n = 0; n < N; n++) {
        uint8_t *base = stage.data();
        base += (16 - ((uintptr_t)base & 15)) & 15;
        uint8_t *rowp = base + ((type == 14) ? phase : 0);
        memcpy(rowp, W + (size_t)n * rb, rb);
        float acc = 0.0f;
        if (type == 12 || type == 13) {
            const U4 *sums4 = (const U4 *)s32.data();
            for (int blk = 0; blk < nblk; blk++) {
                int p = 0, m = 0;
                if (type == 13) acc += block_q45k<true>(rowp + (size_t)blk * 176, aq64.data() + (size_t)blk * 272, sums4[blk], ad[blk], &p, &m);
                else acc += block_q45k<false>(rowp + (size_t)blk * 144, aq64.data() + (size_t)blk * 272, sums4[blk], ad[blk], &p, &m);
                P[(size_t)n * nblk + blk] = p; M[(size_t)n * nblk + blk] = m;
            }
        } else if (type == 14) {
            for (int it = 0; it < K / 128; it++) {
                int p = 0;
                acc += item_q6k(rowp, it, aq128.data(), (const U4 *)s16.data(), ad.data(), &p);
                P[(size_t)n * nblk + (it >> 1)] += p;
            }
        } else return -1;
        dst[n] = acc;
    }
    return 0;
}
