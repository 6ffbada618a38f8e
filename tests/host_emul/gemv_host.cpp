// Host emulation of the GEMV item decoders (cortex.llamacpp_b200/csrc/gemv_items.cuh): the very same
// decode code the CUDA kernel runs per lane, compiled for the CPU so that the bit-twiddling can be checked
// against the oracle without a GPU.  TEST INFRASTRUCTURE.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../cortex.llamacpp_b200/csrc/gemv_items.cuh"

using namespace gemv;

struct Lay { size_t off_d, off_sums, col_bytes; };
static Lay make_lay(int q8k, int64_t K) {
    Lay L; const size_t nd = K / (q8k ? 256 : 32), ns = K / (q8k ? 16 : 32);
    L.off_d = ((size_t)K + 15) & ~(size_t)15;
    L.off_sums = L.off_d + ((nd * 4 + 15) & ~(size_t)15);
    L.col_bytes = L.off_sums + ((ns * 2 + 15) & ~(size_t)15);
    return L;
}

template <int TYPE>
static void run(const uint8_t *W, size_t rb, int N, int K, const uint8_t *act, float *dst, int32_t *P, int32_t *M, int phase) {
    constexpr int ITEM = Traits<TYPE>::ITEM, ASTR = ITEM + 16, Q8K = Traits<TYPE>::Q8K;
    const Lay L = make_lay(Q8K, K);
    const int nitems = num_items<TYPE>(K);
    std::vector<int8_t> aq((size_t)nitems * ASTR + 64, 0);
    for (int e = 0; e < K; e++) aq[(size_t)(e / ITEM) * ASTR + e % ITEM] = (int8_t)act[e];
    ActView A;
    const size_t nd = K / (Q8K ? 256 : 32), ns = K / (Q8K ? 16 : 32);
    std::vector<float> ad(nd + 8, 0.0f); std::vector<int16_t> as(ns + 16, 0);
    memcpy(ad.data(), act + L.off_d, nd * 4); memcpy(as.data(), act + L.off_sums, ns * 2);
    A.q = aq.data(); A.d = ad.data(); A.s = as.data();
    A.q_stride = A.d_stride = A.s_stride = 0;
    const int nblk = K / Traits<TYPE>::BLOCK;
    // emulate a shared-memory stage: row copied to an address with the requested 2-byte phase
    std::vector<uint8_t> stage(rb + 64);
    for (int n = 0; n < N; n++) {
        uint8_t *base = stage.data();
        base += (16 - ((uintptr_t)base & 15)) & 15;
        uint8_t *rowp = base + phase;
        memcpy(rowp, W + (size_t)n * rb, rb);
        float acc[1] = {0.0f};
        DbgSink dbg; dbg.P = P + (size_t)n * nblk; dbg.M = M + (size_t)n * nblk;
        for (int it = 0; it < nitems; it++) {
            ActRegs<TYPE> ar;
            load_act<TYPE>(A, 0, it, ar);
            acc[0] += item_dot<TYPE, true>(rowp, it, K, ar, dbg);
        }
        dst[n] = acc[0];
    }
}

extern "C" int emul_gemv(int type, const uint8_t *W, size_t rb, int N, int K, const uint8_t *act, float *dst, int32_t *P, int32_t *M, int phase) {
    switch (type) {
        case T_Q4_0: run<T_Q4_0>(W, rb, N, K, act, dst, P, M, phase); break;
        case T_Q8_0: run<T_Q8_0>(W, rb, N, K, act, dst, P, M, phase); break;
        case T_Q4_K: run<T_Q4_K>(W, rb, N, K, act, dst, P, M, 0); break;
        case T_Q5_K: run<T_Q5_K>(W, rb, N, K, act, dst, P, M, 0); break;
        case T_Q6_K: run<T_Q6_K>(W, rb, N, K, act, dst, P, M, phase); break;
        default: return -1;
    }
    return 0;
}
