// common.cuh -- internal declarations shared by every translation unit of libggml_b200_kernels.so
// B200 (sm_100a) only.  Replaces the role of ggml-cuda/common.cuh (context, reductions, dp4a) with a
// one-stream-per-context design and Blackwell primitives (mbarrier + cp.async.bulk, PDL).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <unordered_map>

#include "../../include/ggml_b200.h"

// ------------------------------------------------------------------------------------------------
// error handling: nothing throws across the C ABI; failures set a thread-local message
// ------------------------------------------------------------------------------------------------
void b200_set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t err__ = (expr);                                                                 \
        if (err__ != cudaSuccess) {                                                                 \
            b200_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(err__)); \
            return B200_ERR_FAILED;                                                                 \
        }                                                                                           \
    } while (0)

// ------------------------------------------------------------------------------------------------
// block layouts (restated from ggml-common.h:161-328; byte offsets only, no structs shared)
// ------------------------------------------------------------------------------------------------
namespace blk {
constexpr int QK   = 32;
constexpr int QKK  = 256;
constexpr int Q4_0_BYTES = 18;    // half d | qs[16]
constexpr int Q8_0_BYTES = 34;    // half d | int8 qs[32]
constexpr int Q4_K_BYTES = 144;   // half d, half dmin | scales[12] | qs[128]
constexpr int Q5_K_BYTES = 176;   // half d, half dmin | scales[12] | qh[32] | qs[128]
constexpr int Q6_K_BYTES = 210;   // ql[128] | qh[64] | int8 scales[16] | half d
constexpr int Q8_K_BYTES = 292;   // float d | int8 qs[256] | int16 bsums[16]
}  // namespace blk

__host__ __device__ inline int b200_type_block_elems(int type) {
    switch (type) {
        case B200_TYPE_Q4_0: case B200_TYPE_Q8_0: return 32;
        case B200_TYPE_Q4_K: case B200_TYPE_Q5_K: case B200_TYPE_Q6_K: case B200_TYPE_Q8_K: return 256;
        default: return 1;
    }
}
__host__ __device__ inline int b200_type_block_bytes(int type) {
    switch (type) {
        case B200_TYPE_F32: case B200_TYPE_I32: return 4;
        case B200_TYPE_F16: case B200_TYPE_BF16: return 2;
        case B200_TYPE_Q4_0: return 18;
        case B200_TYPE_Q8_0: return 34;
        case B200_TYPE_Q4_K: return 144;
        case B200_TYPE_Q5_K: return 176;
        case B200_TYPE_Q6_K: return 210;
        case B200_TYPE_Q8_K: return 292;
        default: return 0;
    }
}
inline bool b200_type_is_quant(int type) {
    return type == B200_TYPE_Q4_0 || type == B200_TYPE_Q8_0 || type == B200_TYPE_Q4_K || type == B200_TYPE_Q5_K ||
           type == B200_TYPE_Q6_K;
}
inline size_t b200_row_bytes(int type, int64_t k) {
    return (size_t)(k / b200_type_block_elems(type)) * b200_type_block_bytes(type);
}
// activation format the CPU oracle pairs with a weight type (vec_dot_type, ggml-cpu.c:266-341)
inline int b200_act_mode_q8k(int type) { return type == B200_TYPE_Q4_K || type == B200_TYPE_Q5_K || type == B200_TYPE_Q6_K; }

inline bool tensor_is_contiguous(const b200_tensor &t) {
    const int be = b200_type_block_elems(t.type), bb = b200_type_block_bytes(t.type);
    if (t.nb[0] != (uint64_t)bb) return false;
    uint64_t expect = (uint64_t)(t.ne[0] / be) * bb;
    for (int i = 1; i < 4; i++) {
        if (t.ne[i] != 1 && t.nb[i] != expect) return false;
        expect *= t.ne[i];
    }
    return true;
}
inline int64_t tensor_nelements(const b200_tensor &t) { return t.ne[0] * t.ne[1] * t.ne[2] * t.ne[3]; }
inline int64_t tensor_nrows(const b200_tensor &t) { return t.ne[1] * t.ne[2] * t.ne[3]; }

// ------------------------------------------------------------------------------------------------
// activation scratch: the quantised form of src1 that GEMV/GEMM kernels consume.  One column
// (= one token) is laid out as
//     [ qs   : K int8                     ]  (16-byte aligned start)
//     [ d    : K/G floats                 ]  G = 256 (q8_K mode) or 32 (q8_0 mode)
//     [ sums : K/S int16                  ]  S = 16 (q8_K bsums) or 32 (q8_0 mode block sums)
// values are bit-identical to the reference's block_q8_K / block_q8_0 fields (d of q8_0 mode is the
// fp16-rounded scale widened to f32); the split layout is what lets kernels use 128-bit loads.
// ------------------------------------------------------------------------------------------------
struct ActLayout {
    int    q8k;          // 1: q8_K mode, 0: q8_0 mode
    int64_t K;
    size_t off_d, off_sums, col_bytes;
    __host__ __device__ static ActLayout make(int q8k, int64_t K) {
        ActLayout L;
        L.q8k = q8k; L.K = K;
        const size_t nd = (size_t)(K / (q8k ? 256 : 32)), ns = (size_t)(K / (q8k ? 16 : 32));
        L.off_d = ((size_t)K + 15) & ~(size_t)15;
        L.off_sums = L.off_d + ((nd * 4 + 15) & ~(size_t)15);
        L.col_bytes = L.off_sums + ((ns * 2 + 15) & ~(size_t)15);
        return L;
    }
};

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct GraphCache;   // graph.cu
struct b200_comm;    // comm.cu

struct b200_ctx {
    int           device   = 0;
    cudaStream_t  stream   = nullptr;
    int           sm_count = 148;
    size_t        smem_optin = 227 * 1024;
    // growable device scratch (activation quantisation, split-KV partials, MoE routing tables)
    void *        scratch[5]      = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t        scratch_size[5] = {0, 0, 0, 0, 0};
    int64_t       launches = 0;
    int64_t       scratch_gen = 0;          // bumped on every scratch reallocation (captured graphs are dropped with it)
    // options
    int           opt_cuda_graphs = 0;
    int           opt_fusion      = 2;      // 0 off, 1 two-op fusions, 2 + llama layer fusions for decode ubatches
    int           opt_pdl         = 0;
    int           opt_l2_prefetch = 0;      // L2 look-ahead of the next matmul's weights: measured neutral on B200 (profiles/r1_gemv_diag.md), off
    int           opt_fa_exact    = 0;      // 1: f16-cache flash attention reproduces the CPU's fp16 V accumulator cell by cell (parity mode, serial over n_kv)
    int           opt_debug_skip  = 0;      // timing experiments only: bit 0 flash_attn, 1 rope+store, 2 GEMV are not launched
    GraphCache *  graph_cache = nullptr;
    void *        fattn_counters = nullptr;   // split-arrival counters of the fused flash-attention combine (fattn.cu)
    b200_comm *   comm = nullptr;       // tensor-parallel communicator (comm.cu); NULL = single GPU
    bool          capturing = false;
    void *        prof_buf = nullptr;   // debug: per-CTA timestamps (b200_debug_set_prof)
    int           prof_launch = 0;

    void *get_scratch(int slot, size_t size);   // grows (sync + realloc) when too small; nullptr on OOM
};

enum { SCRATCH_ACT = 0, SCRATCH_FATTN = 1, SCRATCH_MOE = 2, SCRATCH_MISC = 3, SCRATCH_FUSE = 4, SCRATCH_COUNT = 5 };

// op entry points (each in its own .cu); all asynchronous on ctx->stream
int op_mul_mat(b200_ctx *ctx, const b200_op *op);
int op_mul_mat_id(b200_ctx *ctx, const b200_op *op);
int op_flash_attn_ext(b200_ctx *ctx, const b200_op *op);
int op_glue(b200_ctx *ctx, const b200_op *op);          // everything else
bool supports_mul_mat(const b200_op *op);
bool supports_mul_mat_id(const b200_op *op);
bool supports_flash_attn_ext(const b200_op *op);
bool supports_glue(const b200_op *op);
int op_allreduce(b200_ctx *ctx, const b200_op *op);     // comm.cu
bool supports_allreduce(const b200_op *op);

// fused ROPE(q) + ROPE(k) + KV-store of k and v for a decode ubatch (glue.cu); q/k/v are contiguous [D, heads, T] f32
struct RopeStoreDesc {
    const float *q, *k, *v;
    const int32_t *pos;
    const float *freq_factors;          // optional [D/2]
    float *q_out; uint64_t q_out_nb1, q_out_nb2;     // roped q, byte strides of head and token
    void *k_dst, *v_dst;                // cache rows of token 0 of this ubatch (layout [T][Hkv][D] in the cache type)
    void *const *k_dst_ind, *const *v_dst_ind;       // when non-NULL the destinations are read from device memory (CUDA-graph replay)
    int D, H, Hkv, T, kv_type;
    int32_t rope_params[16];
    int use_pdl;
};
int launch_rope_store(b200_ctx *ctx, const RopeStoreDesc &d);

// quantise f32 activations [K, ncols] (column byte stride nb1) into the scratch layout above
int launch_quantize_act(b200_ctx *ctx, int q8k, const float *x, size_t x_col_stride_bytes, int64_t K, int64_t ncols,
                        uint8_t *scratch);
// decode GEMV over pre-quantised activations (gemv.cu)
int launch_gemv(b200_ctx *ctx, int type, const uint8_t *W, size_t row_bytes, int64_t N, int64_t K, const uint8_t *act,
                int ncols, float *dst, size_t dst_col_stride_elems, bool w_const);

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float warp_reduce_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_reduce_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_reduce_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float half_bits_to_float(uint32_t h16) {
    return __half2float(__ushort_as_half((unsigned short)(h16 & 0xffffu)));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier (Hopper+/Blackwell producer-consumer pipeline) ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D bulk async copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// programmatic dependent launch
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

#endif  // __CUDACC__
