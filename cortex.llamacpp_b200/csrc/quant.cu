// quant.cu -- activation quantisers.
//
// Replaces quantize_q8_1 (ggml-cuda/quantize.cu:4-38) with quantisers that reproduce the CPU oracle's
// activation formats bit for bit, because parity is defined against the CPU backend:
//   q8_K mode (K-quant weights): quantize_row_q8_K_ref, ggml-quants.c:2479-2513
//       iscale = -127/max (max = signed value of the FIRST max-|x| element), q = min(127, RNE(iscale*x)),
//       d = 1/iscale (f32), bsums per 16
//   q8_0 mode (Q4_0/Q8_0 weights, quantised-K attention): quantize_row_q8_0 AVX2 path,
//       ggml-cpu-quants.c:808-860:  d = amax/127 -> fp16, id = 127/amax, q = RNE(x*id)
// Output goes to the split "activation scratch" layout described in common.cuh (ActLayout).
#include "common.cuh"
#include "quant_warp.cuh"

namespace {

// one warp per 256 elements; lane owns 8 consecutive elements
__global__ void __launch_bounds__(128) b200_quantize_q8k_kernel(const float *__restrict__ x, size_t x_col_stride, int64_t K,
                                                           int64_t ncols, uint8_t *__restrict__ out, ActLayout L) {
    const int lane = threadIdx.x & 31;
    const int64_t nblk = K / 256;
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= nblk * ncols) return;
    const int64_t col = gw / nblk, b = gw % nblk;
    const float *xp = (const float *)((const char *)x + col * x_col_stride) + b * 256 + lane * 8;
    uint8_t *oc = out + col * L.col_bytes;

    float v[8];
    {
        const float4 a = *(const float4 *)xp, c = *(const float4 *)(xp + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    }
    uint2 qp; float d; int pair;
    warp_quant_q8k(v, lane, qp, d, pair);
    *(uint2 *)(oc + b * 256 + lane * 8) = qp;
    if ((lane & 1) == 0) ((int16_t *)(oc + L.off_sums))[b * 16 + (lane >> 1)] = (int16_t)pair;
    if (lane == 0) ((float *)(oc + L.off_d))[b] = d;
}

// one warp per 256 elements = 8 blocks of 32; 4 lanes per block
__global__ void __launch_bounds__(128) b200_quantize_q80_kernel(const float *__restrict__ x, size_t x_col_stride, int64_t K,
                                                           int64_t ncols, uint8_t *__restrict__ out, ActLayout L) {
    const int lane = threadIdx.x & 31;
    const int64_t nchunk = (K + 255) / 256;
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= nchunk * ncols) return;
    const int64_t col = gw / nchunk, c = gw % nchunk;
    const int64_t e0 = c * 256 + lane * 8;
    const bool valid = e0 < K;                 // K % 32 == 0 so a block is entirely valid or not
    const float *xp = (const float *)((const char *)x + col * x_col_stride) + e0;
    uint8_t *oc = out + col * L.col_bytes;

    float v[8];
    if (valid) {
        const float4 a = *(const float4 *)xp, b = *(const float4 *)(xp + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = 0.0f;
    }
    uint2 qp; float d16; int lsum;
    warp_quant_q80(v, qp, d16, lsum);
    if (!valid) return;
    *(uint2 *)(oc + e0) = qp;
    if ((lane & 3) == 0) {
        const int64_t bi = e0 / 32;
        ((float *)(oc + L.off_d))[bi] = d16;
        ((int16_t *)(oc + L.off_sums))[bi] = (int16_t)lsum;
    }
}

// Producer of a small-batch matmul's activations in ONE launch: (x | rms_norm(x) * w | silu(gate) * up) -> quantised scratch.
// mode = ACT_F32 / ACT_F32_NORM / ACT_F32_SWIGLU of gemv.h; same arithmetic as the GEMV prologue (gemv.cu) and the stand-alone
// glue kernels: rms_norm with the sum of squares in double, ggml_v_expf silu.  grid = (columns, K splits), 8 warps per CTA.
__global__ void __launch_bounds__(256) b200_act_prologue_kernel(int mode, const float *__restrict__ x, size_t x_stride, const float *__restrict__ x2, float eps,
                                                              int K, int q8k, uint8_t *__restrict__ out, ActLayout L, int use_pdl) {
    __shared__ double sred[8];
    if (use_pdl) { pdl_trigger(); pdl_wait(); }
    const int col = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *xp = (const float *)((const char *)x + (size_t)col * x_stride);
    const float *up = mode == 3 ? (const float *)((const char *)x2 + (size_t)col * x_stride) : nullptr;
    float norm_scale = 1.0f;
    if (mode == 2) {
        double ss = 0.0;
        for (int i = threadIdx.x * 4; i < K; i += 1024) {
            const float4 v = *(const float4 *)(xp + i);
            ss += (double)__fmul_rn(v.x, v.x); ss += (double)__fmul_rn(v.y, v.y); ss += (double)__fmul_rn(v.z, v.z); ss += (double)__fmul_rn(v.w, v.w);
        }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) sred[warp] = ss;
        __syncthreads();
        double t = 0.0;
        for (int i = 0; i < 8; i++) t += sred[i];
        const float mean = (float)(t / (double)K);
        norm_scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
    }
    uint8_t *oc = out + (size_t)col * L.col_bytes;
    const int nchunk = (K + 255) / 256;
    for (int b = blockIdx.y * 8 + warp; b < nchunk; b += 8 * gridDim.y) {
        const int e0 = b * 256 + lane * 8;
        const bool valid = e0 < K;
        float v[8];
        if (valid) {
            const float4 a = *(const float4 *)(xp + e0), c = *(const float4 *)(xp + e0 + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
            if (mode >= 2) {
                const float *wp = mode == 2 ? x2 : up;
                const float4 wa = *(const float4 *)(wp + e0), wc = *(const float4 *)(wp + e0 + 4);
                const float w[8] = {wa.x, wa.y, wa.z, wa.w, wc.x, wc.y, wc.z, wc.w};
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = mode == 2 ? __fmul_rn(__fmul_rn(v[j], norm_scale), w[j]) : __fmul_rn(ggml_silu_lane(v[j]), w[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = 0.0f;
        }
        uint2 qp;
        if (q8k) {
            float d; int pair;
            warp_quant_q8k(v, lane, qp, d, pair);
            *(uint2 *)(oc + e0) = qp;
            if ((lane & 1) == 0) ((int16_t *)(oc + L.off_sums))[b * 16 + (lane >> 1)] = (int16_t)pair;
            if (lane == 0) ((float *)(oc + L.off_d))[b] = d;
        } else {
            float d16; int bsum;
            warp_quant_q80(v, qp, d16, bsum);
            if (valid) {
                *(uint2 *)(oc + e0) = qp;
                if ((lane & 3) == 0) { ((float *)(oc + L.off_d))[e0 >> 5] = d16; ((int16_t *)(oc + L.off_sums))[e0 >> 5] = (int16_t)bsum; }
            }
        }
    }
}

// scratch layout -> the reference's canonical block bytes (test hook only)
__global__ void b200_repack_q8k_kernel(const uint8_t *__restrict__ in, ActLayout L, int64_t ncols, uint8_t *__restrict__ blocks) {
    const int64_t nblk = L.K / 256;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblk * ncols) return;
    const int64_t col = i / nblk, b = i % nblk;
    const uint8_t *ic = in + col * L.col_bytes;
    uint8_t *o = blocks + i * 292;
    *(float *)o = ((const float *)(ic + L.off_d))[b];
    for (int j = 0; j < 256; j++) o[4 + j] = ic[b * 256 + j];
    for (int j = 0; j < 16; j++) *(int16_t *)(o + 260 + 2 * j) = ((const int16_t *)(ic + L.off_sums))[b * 16 + j];
}
__global__ void b200_repack_q80_kernel(const uint8_t *__restrict__ in, ActLayout L, int64_t ncols, uint8_t *__restrict__ blocks) {
    const int64_t nblk = L.K / 32;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblk * ncols) return;
    const int64_t col = i / nblk, b = i % nblk;
    const uint8_t *ic = in + col * L.col_bytes;
    uint8_t *o = blocks + i * 34;
    *(__half *)o = __float2half_rn(((const float *)(ic + L.off_d))[b]);   // exact: value is already an fp16
    for (int j = 0; j < 32; j++) o[2 + j] = ic[b * 32 + j];
}

}  // namespace

int launch_quantize_act(b200_ctx *ctx, int q8k, const float *x, size_t x_col_stride, int64_t K, int64_t ncols, uint8_t *scratch) {
    const ActLayout L = ActLayout::make(q8k, K);
    const int64_t warps = (q8k ? K / 256 : (K + 255) / 256) * ncols;
    if (warps == 0) return B200_OK;
    const unsigned grid = (unsigned)((warps + 3) / 4);
    if (q8k) b200_quantize_q8k_kernel<<<grid, 128, 0, ctx->stream>>>(x, x_col_stride, K, ncols, scratch, L);
    else     b200_quantize_q80_kernel<<<grid, 128, 0, ctx->stream>>>(x, x_col_stride, K, ncols, scratch, L);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

// mode: 1 plain f32, 2 rms_norm(x) * x2, 3 silu(x) * x2 (x2 strided like x)
int launch_act_prologue(b200_ctx *ctx, int mode, const float *x, size_t x_stride_bytes, const float *x2, float eps, int64_t K, int ncols, int q8k, uint8_t *scratch) {
    const ActLayout L = ActLayout::make(q8k, K);
    const int nchunk = (int)((K + 255) / 256);
    int nsplit = (nchunk + 7) / 8;
    const int cap = (2 * ctx->sm_count + ncols - 1) / ncols;
    if (nsplit > cap) nsplit = cap;
    if (nsplit < 1) nsplit = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ncols, (unsigned)nsplit); cfg.blockDim = dim3(256); cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = ctx->opt_pdl ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, b200_act_prologue_kernel, mode, x, x_stride_bytes, x2, eps, (int)K, q8k, scratch, L, ctx->opt_pdl));
    ctx->launches++;
    return B200_OK;
}

extern "C" int b200_quantize_act(b200_ctx *ctx, int32_t act_type, const float *x, void *blocks, int64_t K, int64_t rows) {
    if (!ctx) return B200_ERR_FAILED;
    const int q8k = act_type == B200_TYPE_Q8_K;
    if (!q8k && act_type != B200_TYPE_Q8_0) { b200_set_error("quantize_act: type %d", act_type); return B200_ERR_UNSUPPORTED; }
    if (K % (q8k ? 256 : 32) != 0) { b200_set_error("quantize_act: K=%lld", (long long)K); return B200_ERR_UNSUPPORTED; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const ActLayout L = ActLayout::make(q8k, K);
    uint8_t *s = (uint8_t *)ctx->get_scratch(SCRATCH_ACT, L.col_bytes * (size_t)rows);
    if (!s) return B200_ERR_ALLOC;
    int rc = launch_quantize_act(ctx, q8k, x, (size_t)K * 4, K, rows, s);
    if (rc) return rc;
    const int64_t n = (K / (q8k ? 256 : 32)) * rows;
    if (q8k) b200_repack_q8k_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(s, L, rows, (uint8_t *)blocks);
    else     b200_repack_q80_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(s, L, rows, (uint8_t *)blocks);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}
