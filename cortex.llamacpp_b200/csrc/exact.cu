// exact.cu -- "cpu-exact" parity mode of MUL_MAT (option cpu_exact / GGML_B200_CPU_EXACT=1).
//
// Why it exists.  Every kernel of this backend reproduces the CPU backend's INTEGER arithmetic bit for bit (per-block
// sums), but the fast kernels add the per-block float terms in their own order (block per lane, shuffle trees, split-K), so
// a matmul output differs from the CPU's in its last bit or two.  The reference pipeline amplifies such differences: the next
// op quantises its input to int8 (q8_0 / q8_K), where a 1e-7 change flips a rounding now and then, each flip is worth 1/127
// of a block maximum, and the error grows roughly as a square root per matmul until it saturates around 1e-2 of the logits
// after a few layers (measured node by node: profiles/r2_node_divergence.txt).  Two builds of the reference itself (AVX2 vs
// AVX512, llamafile sgemm on/off, repacked weights) differ in exactly the same way.  To show that summation order is the
// ONLY difference, this mode computes every dst element in the order of the reference's AVX2 build:
//   ggml_vec_dot_{q4_0,q8_0}_q8_0 (ggml-cpu-quants.c:2273-2296, 3935-3952), ggml_vec_dot_{q4_K,q5_K,q6_K}_q8_K
//   (:6776-6837, 7413-7490, 8405-8481): an 8-lane int32 vector per weight block (lane l = bytes 4l..4l+3 of every
//   32-element group, times the group's sub-scale), one FMA per block into an 8-lane f32 accumulator, hsum_float_8 (:49-55)
//   at the end; q4_K mins through a 4-lane FMA accumulator, q5_K mins through a scalar float.
// Eight threads play the eight SIMD lanes of one (row, column) dot product.  It is a verification mode (uncoalesced,
// serial over K), not a fast path: tests/test_gpu_reference_parity.py runs the model-level comparisons in both modes.
#include "common.cuh"

namespace {

__device__ __forceinline__ int dot4(uint32_t w_signed_bytes, uint32_t a) { return __dp4a((int)w_signed_bytes, (int)a, 0); }
__device__ __forceinline__ uint32_t ld4_unaligned(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
// get_scale_min_k4 (ggml-quants.c:631-639)
__device__ __forceinline__ void scale_min_k4(const uint8_t *s, int j, int &sc, int &mn) {
    if (j < 4) { sc = s[j] & 63; mn = s[j + 4] & 63; }
    else { sc = (s[j + 4] & 0x0f) | ((s[j - 4] >> 6) << 4); mn = (s[j + 4] >> 4) | ((s[j] >> 6) << 4); }
}

template <int TYPE>
__global__ void __launch_bounds__(256) b200_mul_mat_exact_kernel(const uint8_t *W, size_t rb, int64_t N, int64_t K, const uint8_t *act,
                                                                 ActLayout L, int64_t ncols, float *dst, size_t dst_stride, ExactMoe moe) {
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int l = threadIdx.x & 7;
    if (gid >= N * ncols) return;                         // whole 8-thread groups leave together (256 % 8 == 0)
    const int64_t row = gid % N;
    int64_t col = gid / N;
    size_t dst_off = (size_t)col * dst_stride;
    if (moe.ids) {                                        // MUL_MAT_ID: column = (token, slot) pair, expert read on the device
        const int64_t t = col / moe.n_used, sl = col % moe.n_used;
        W += (size_t)(*(const int32_t *)(moe.ids + sl * moe.ids_nb0 + t * moe.ids_nb1)) * moe.expert_stride;
        dst_off = (size_t)sl * moe.d_nb1 + (size_t)t * moe.d_nb2;
        if (moe.b_ne1 == 1) col = t;
    }
    const uint8_t *wrow = W + (size_t)row * rb;
    const uint8_t *acol = act + (size_t)col * L.col_bytes;
    const int8_t *aq = (const int8_t *)acol;
    const float *ad = (const float *)(acol + L.off_d);
    const int16_t *as = (const int16_t *)(acol + L.off_sums);
    float acc = 0.0f, acc_m = 0.0f;                       // lane l of the 8-lane accumulator; lane l (< 4) of the q4_K mins accumulator / q5_K scalar (lane 0)
    constexpr int BE = (TYPE == B200_TYPE_Q4_0 || TYPE == B200_TYPE_Q8_0) ? 32 : 256;
    const int64_t nb = K / BE;
    for (int64_t b = 0; b < nb; b++) {
        int Lb = 0;
        float d;
        if (TYPE == B200_TYPE_Q4_0) {
            const uint8_t *blk = wrow + b * 18;
            const uint32_t raw = ld4_unaligned(blk + 2 + 4 * (l & 3));
            const uint32_t nib = l < 4 ? (raw & 0x0f0f0f0fu) : ((raw >> 4) & 0x0f0f0f0fu);
            const uint32_t a = *(const uint32_t *)(aq + 32 * b + 4 * l);
            // (q - 8) . a = q . a - 8 * sum(a): keep it as signed bytes instead (q - 8 in [-8, 7])
            const uint32_t q = __vsub4(nib, 0x08080808u);
            Lb = dot4(q, a);
            d = __fmul_rn(__half2float(__ushort_as_half((unsigned short)(blk[0] | (blk[1] << 8)))), ad[b]);
        } else if (TYPE == B200_TYPE_Q8_0) {
            const uint8_t *blk = wrow + b * 34;
            Lb = dot4(ld4_unaligned(blk + 2 + 4 * l), *(const uint32_t *)(aq + 32 * b + 4 * l));
            d = __fmul_rn(__half2float(__ushort_as_half((unsigned short)(blk[0] | (blk[1] << 8)))), ad[b]);
        } else if (TYPE == B200_TYPE_Q4_K || TYPE == B200_TYPE_Q5_K) {
            const uint8_t *blk = wrow + b * (TYPE == B200_TYPE_Q4_K ? 144 : 176);
            const uint8_t *qs = blk + (TYPE == B200_TYPE_Q4_K ? 16 : 48);
            const uint32_t hb = TYPE == B200_TYPE_Q5_K ? *(const uint32_t *)(blk + 16 + 4 * l) : 0u;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                int sc, mn;
                scale_min_k4(blk + 4, j, sc, mn);
                const uint32_t raw = *(const uint32_t *)(qs + 32 * (j >> 1) + 4 * l);
                uint32_t q = (j & 1) ? ((raw >> 4) & 0x0f0f0f0fu) : (raw & 0x0f0f0f0fu);
                if (TYPE == B200_TYPE_Q5_K) q |= ((hb >> j) & 0x01010101u) << 4;
                Lb += sc * dot4(q, *(const uint32_t *)(aq + 256 * b + 32 * j + 4 * l));
            }
            const float da = ad[b];
            d = __fmul_rn(da, __half2float(*(const __half *)blk));
            const float dmin = __fmul_rn(-da, __half2float(*(const __half *)(blk + 2)));
            const int16_t *bs = as + 16 * b;
            if (TYPE == B200_TYPE_Q4_K) {
                if (l < 4) {
                    int sc, m0, m1;
                    scale_min_k4(blk + 4, 2 * l, sc, m0);
                    scale_min_k4(blk + 4, 2 * l + 1, sc, m1);
                    const int s0 = (int16_t)(bs[4 * l] + bs[4 * l + 1]), s1 = (int16_t)(bs[4 * l + 2] + bs[4 * l + 3]);
                    acc_m = __fmaf_rn(dmin, (float)(m0 * s0 + m1 * s1), acc_m);
                }
            } else if (l == 0) {
                int prod[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    int sc, m0, m1;
                    scale_min_k4(blk + 4, 2 * k, sc, m0);
                    scale_min_k4(blk + 4, 2 * k + 1, sc, m1);
                    const int s0 = (int16_t)(bs[4 * k] + bs[4 * k + 1]), s1 = (int16_t)(bs[4 * k + 2] + bs[4 * k + 3]);
                    prod[k] = m0 * s0 + m1 * s1;
                }
                acc_m = __fadd_rn(acc_m, __fmul_rn(dmin, (float)((prod[0] + prod[1]) + (prod[2] + prod[3]))));
            }
        } else {   // Q6_K: ql[128] | qh[64] | int8 scales[16] | half d, blocks only 2-byte aligned
            const uint8_t *blk = wrow + b * 210;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t qh = ld4_unaligned(blk + 128 + 32 * h + 4 * l);
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const uint32_t raw = ld4_unaligned(blk + 64 * h + 32 * (t & 1) + 4 * l);
                    const uint32_t lo = t < 2 ? (raw & 0x0f0f0f0fu) : ((raw >> 4) & 0x0f0f0f0fu);
                    const uint32_t hi = ((qh >> (2 * t)) & 0x03030303u) << 4;
                    const uint32_t q = __vsub4(lo | hi, 0x20202020u);                      // q - 32 in [-32, 31]
                    const int scale = (int)(int8_t)blk[192 + 8 * h + 2 * t + (l >> 2)];
                    Lb += scale * dot4(q, *(const uint32_t *)(aq + 256 * b + 128 * h + 32 * t + 4 * l));
                }
            }
            d = __fmul_rn(ad[b], __half2float(__ushort_as_half((unsigned short)(blk[208] | (blk[209] << 8)))));
        }
        acc = __fmaf_rn(d, (float)Lb, acc);
    }
    // hsum_float_8: ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)); the 8 lanes of a group are 8 consecutive lanes of the warp
    const unsigned gm = 0xffu << (threadIdx.x & 24);
    float v = __fadd_rn(acc, __shfl_down_sync(gm, acc, 4, 8));
    v = __fadd_rn(v, __shfl_down_sync(gm, v, 2, 8));
    v = __fadd_rn(v, __shfl_down_sync(gm, v, 1, 8));
    if (TYPE == B200_TYPE_Q4_K) {
        float m = __fadd_rn(acc_m, __shfl_down_sync(gm, acc_m, 2, 8));
        m = __fadd_rn(m, __shfl_down_sync(gm, m, 1, 8));
        v = __fadd_rn(v, m);
    } else if (TYPE == B200_TYPE_Q5_K) {
        v = __fadd_rn(v, acc_m);
    }
    if (l == 0) dst[dst_off + row] = v;
}

}  // namespace

int launch_mul_mat_exact(b200_ctx *ctx, int type, const uint8_t *W, size_t rb, int64_t N, int64_t K, const uint8_t *act, int64_t ncols,
                         float *dst, size_t dst_stride, const ExactMoe *moep) {
    ExactMoe moe = {};
    if (moep) moe = *moep;
    const ActLayout L = ActLayout::make(b200_act_mode_q8k(type), K);
    const int64_t groups = N * ncols;
    if (groups == 0) return B200_OK;
    const unsigned grid = (unsigned)((groups * 8 + 255) / 256);
    switch (type) {
        case B200_TYPE_Q4_0: b200_mul_mat_exact_kernel<B200_TYPE_Q4_0><<<grid, 256, 0, ctx->stream>>>(W, rb, N, K, act, L, ncols, dst, dst_stride, moe); break;
        case B200_TYPE_Q8_0: b200_mul_mat_exact_kernel<B200_TYPE_Q8_0><<<grid, 256, 0, ctx->stream>>>(W, rb, N, K, act, L, ncols, dst, dst_stride, moe); break;
        case B200_TYPE_Q4_K: b200_mul_mat_exact_kernel<B200_TYPE_Q4_K><<<grid, 256, 0, ctx->stream>>>(W, rb, N, K, act, L, ncols, dst, dst_stride, moe); break;
        case B200_TYPE_Q5_K: b200_mul_mat_exact_kernel<B200_TYPE_Q5_K><<<grid, 256, 0, ctx->stream>>>(W, rb, N, K, act, L, ncols, dst, dst_stride, moe); break;
        case B200_TYPE_Q6_K: b200_mul_mat_exact_kernel<B200_TYPE_Q6_K><<<grid, 256, 0, ctx->stream>>>(W, rb, N, K, act, L, ncols, dst, dst_stride, moe); break;
        default: b200_set_error("mul_mat exact: type %d", type); return B200_ERR_UNSUPPORTED;
    }
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}
