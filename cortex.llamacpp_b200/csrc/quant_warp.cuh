// quant_warp.cuh -- warp-level activation quantisers shared by quant.cu (stand-alone kernels) and the
// fused GEMV prologue (gemv.cu).  Bit-exact restatements of the CPU oracle's quantisers:
//   q8_K: quantize_row_q8_K_ref (ggml-quants.c:2479-2513)      q8_0: quantize_row_q8_0 AVX2 path (ggml-cpu-quants.c:808-860)
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

// One warp quantises 256 consecutive floats; lane owns v[0..7] = x[8*lane .. 8*lane+7].
// Returns the 8 int8 quants packed in a uint2, the block scale d (all lanes) and, in EVEN lanes, the
// bsum of the 16-element group lane/2.
__device__ __forceinline__ void warp_quant_q8k(const float (&v)[8], int lane, uint2 &qpack, float &d, int &pairsum) {
    float amax = 0.0f, mx = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float a = fabsf(v[j]);
        if (a > amax) { amax = a; mx = v[j]; }
    }
    // block maximum of |x| with the reference's first-occurrence tie break (equal |x| -> lower index wins: the reference scans in order).
    // Non-negative floats order like their bit patterns, so one integer REDUX finds the maximum, a ballot finds the lanes that hold it and
    // the lowest of them supplies the signed value (was: a 5-step tree of 3 shuffles each).  NaN never occurs in activations that matter;
    // a NaN |x| compares false in `a > amax` above exactly as before and is ignored by both versions.
    {
        const unsigned top = __reduce_max_sync(0xffffffffu, __float_as_uint(amax));
        const unsigned holders = __ballot_sync(0xffffffffu, __float_as_uint(amax) == top);
        mx = __shfl_sync(0xffffffffu, mx, __ffs(holders) - 1);
        amax = __uint_as_float(top);
    }
    int8_t q[8];
    int lsum = 0;
    d = 0.0f;
    if (amax != 0.0f) {
        const float iscale = __fdiv_rn(-127.0f, mx);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int t = __float2int_rn(__fmul_rn(iscale, v[j]));
            t = t > 127 ? 127 : t;
            q[j] = (int8_t)t;
            lsum += t;
        }
        d = __fdiv_rn(1.0f, iscale);
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) q[j] = 0;
    }
    qpack = *(const uint2 *)q;
    pairsum = lsum + __shfl_xor_sync(0xffffffffu, lsum, 1);
}

// One warp quantises 8 consecutive 32-element blocks; lane owns 8 elements, 4 lanes per block.
// Returns packed quants, the block scale as the f32 value of its fp16 rounding, and the block's quant sum.
__device__ __forceinline__ void warp_quant_q80(const float (&v)[8], uint2 &qpack, float &d_f16, int &blocksum) {
    float amax = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; j++) amax = fmaxf(amax, fabsf(v[j]));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    const float d  = __fdiv_rn(amax, 127.0f);
    const float id = amax != 0.0f ? __fdiv_rn(127.0f, amax) : 0.0f;
    int8_t q[8];
    int lsum = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int t = __float2int_rn(__fmul_rn(v[j], id));
        q[j] = (int8_t)t;
        lsum += t;
    }
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
    qpack = *(const uint2 *)q;
    d_f16 = __half2float(__float2half_rn(d));
    blocksum = lsum;
}
