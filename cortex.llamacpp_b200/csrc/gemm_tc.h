// gemm_tc.h -- host interface of the tcgen05 kind::f16 K-quant GEMM (gemm_tc.cu)
#pragma once
#include "common.cuh"
bool gemm_tc_supported(int type, int64_t N, int64_t K, int64_t M);
int gemm_tc_run(b200_ctx *ctx, int type, const uint8_t *W, size_t rb, int64_t N, int64_t K, const float *x, size_t x_stride, int64_t M,
                float *dst, size_t dst_stride);
int gemm_tc_run_grouped(b200_ctx *ctx, int type, const uint8_t *W, size_t rb, int64_t N, int64_t K, const float *x, size_t x_stride, const MmGroupDesc &g, float *dst);
