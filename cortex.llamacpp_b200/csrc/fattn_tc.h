// fattn_tc.h -- host interface of the tcgen05 flash-attention kernel for prompt-sized query blocks (fattn_tc.cu)
#pragma once
#include "common.cuh"
struct FaTcArgs {
    const char *q; uint64_t q_nb1, q_nb2;          // f32 [128, n_q, H]
    const char *k; uint64_t k_nb1, k_nb2;          // f16 / q8_0 / q4_0 [128, n_kv, Hkv]
    const char *v; uint64_t v_nb1, v_nb2;
    const char *mask; uint64_t m_nb1;              // f16 [n_kv, >= n_q] or NULL
    float *dst;                                    // f32 [128, H, n_q], contiguous
    int n_q, n_kv, H, gq;
    float scale;
};
int fattn_tc_launch(b200_ctx *ctx, const FaTcArgs &p, int kv_type);
