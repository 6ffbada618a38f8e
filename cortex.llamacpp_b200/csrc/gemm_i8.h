// gemm_i8.h -- host interface of the tcgen05 int8 prefill GEMM (gemm_i8.cu)
#pragma once
#include "common.cuh"
bool gemm_i8_supported(int type, int64_t N, int64_t K, int64_t M);
int gemm_i8_run(b200_ctx *ctx, int type, const uint8_t *W, size_t rb, int64_t N, int64_t K, const float *x, size_t x_stride, int64_t M,
                float *dst, size_t dst_stride);
