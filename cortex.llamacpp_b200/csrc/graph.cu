// graph.cu -- op dispatch and graph execution behind b200_graph_compute().
//
// Replaces ggml_cuda_compute_forward (ggml-cuda.cu:2100-2332), ggml_backend_cuda_device_supports_op
// (:2959-3258) and the CUDA-graph capture/update logic of ggml_backend_cuda_graph_compute (:2432-2788).
// The caller (ggml-b200.cpp or a test) hands over a flat list of b200_op; we
//   1. optionally fuse adjacent ops (RMS_NORM+MUL, SILU+MUL) when the intermediate is not observable,
//   2. run them on the context's stream, or
//   3. when cuda_graphs is on and the very same op list (pointers, shapes, params) was seen before,
//      replay the captured cudaGraphExec instead of re-launching ~10 kernels per layer.
#include "common.cuh"
#include <string.h>

struct GraphCacheEntry {
    std::vector<b200_op> ops;
    cudaGraphExec_t exec = nullptr;
    int hits = 0;
};
struct GraphCache {
    std::vector<GraphCacheEntry> entries;
};

void graph_cache_free(b200_ctx *ctx) {
    if (!ctx->graph_cache) return;
    for (auto &e : ctx->graph_cache->entries) if (e.exec) cudaGraphExecDestroy(e.exec);
    delete ctx->graph_cache;
    ctx->graph_cache = nullptr;
}

static int dispatch(b200_ctx *ctx, const b200_op *op) {
    switch (op->op) {
        case B200_OP_NONE: return B200_OK;
        case B200_OP_MUL_MAT: return op_mul_mat(ctx, op);
        case B200_OP_MUL_MAT_ID: return op_mul_mat_id(ctx, op);
        case B200_OP_FLASH_ATTN_EXT: return op_flash_attn_ext(ctx, op);
        default: return op_glue(ctx, op);
    }
}

extern "C" int b200_supports_op(int device, const b200_op *op) {
    (void)device;
    if (!op) return 0;
    switch (op->op) {
        case B200_OP_NONE: return 1;
        case B200_OP_MUL_MAT: return supports_mul_mat(op) ? 1 : 0;
        case B200_OP_MUL_MAT_ID: return supports_mul_mat_id(op) ? 1 : 0;
        case B200_OP_FLASH_ATTN_EXT: return supports_flash_attn_ext(op) ? 1 : 0;
        default: return supports_glue(op) ? 1 : 0;
    }
}

extern "C" int b200_op_compute(b200_ctx *ctx, const b200_op *op) {
    if (!ctx || !op) return B200_ERR_FAILED;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (!b200_supports_op(ctx->device, op)) { b200_set_error("op %d not supported for these operands", op->op); return B200_ERR_UNSUPPORTED; }
    return dispatch(ctx, op);
}

// ---------------------------------------------------------------------------------------------- fusion
static bool same_tensor(const b200_tensor &a, const b200_tensor &b) {
    return a.data == b.data && a.type == b.type && !memcmp(a.ne, b.ne, sizeof(a.ne)) && !memcmp(a.nb, b.nb, sizeof(a.nb));
}
// is tensor t read by any op in [from, n) other than `except`?
static bool read_later(const b200_op *ops, int n, int from, int except, const b200_tensor &t) {
    for (int i = from; i < n; i++) {
        if (i == except) continue;
        for (int s = 0; s < ops[i].n_src; s++) if (ops[i].src[s].data == t.data) return true;
    }
    return false;
}

// returns the list actually executed
static void fuse(const b200_op *ops, int n, std::vector<b200_op> &out) {
    out.clear();
    out.reserve(n);
    for (int i = 0; i < n; i++) {
        const b200_op &a = ops[i];
        if (i + 1 < n) {
            const b200_op &b = ops[i + 1];
            // RMS_NORM -> MUL(norm, weight): y = rms_norm(x) * w
            if (a.op == B200_OP_RMS_NORM && b.op == B200_OP_MUL && same_tensor(b.src[0], a.dst) && b.src[1].type == B200_TYPE_F32 &&
                b.src[1].nb[0] == 4 && b.src[1].ne[0] == a.dst.ne[0] &&
                (b.dst.data == a.dst.data || !read_later(ops, n, i + 2, -1, a.dst))) {
                b200_op f = a;
                f.op = B200_OP_RMS_NORM_MUL;
                f.n_src = 2;
                f.src[1] = b.src[1];
                f.dst = b.dst;
                if (supports_glue(&f)) { out.push_back(f); i++; continue; }
            }
            // SILU(gate) -> MUL(silu, up)
            if (a.op == B200_OP_SILU && b.op == B200_OP_MUL && same_tensor(b.src[0], a.dst) &&
                (b.dst.data == a.dst.data || !read_later(ops, n, i + 2, -1, a.dst))) {
                b200_op f = a;
                f.op = B200_OP_SWIGLU_FUSED;
                f.n_src = 2;
                f.src[1] = b.src[1];
                f.dst = b.dst;
                if (supports_glue(&f)) { out.push_back(f); i++; continue; }
            }
        }
        out.push_back(a);
    }
}

static int run_list(b200_ctx *ctx, const std::vector<b200_op> &ops) {
    for (const b200_op &op : ops) {
        int rc = dispatch(ctx, &op);
        if (rc) return rc;
    }
    return B200_OK;
}

static bool ops_equal(const std::vector<b200_op> &a, const b200_op *b, int n) {
    if ((int)a.size() != n) return false;
    return memcmp(a.data(), b, sizeof(b200_op) * (size_t)n) == 0;
}

extern "C" int b200_graph_compute(b200_ctx *ctx, const b200_op *ops, int n_ops) {
    if (!ctx || (!ops && n_ops)) return B200_ERR_FAILED;
    CUDA_TRY(cudaSetDevice(ctx->device));
    for (int i = 0; i < n_ops; i++)
        if (!b200_supports_op(ctx->device, &ops[i])) {
            b200_set_error("graph op %d (id %d) not supported", i, ops[i].op);
            return B200_ERR_UNSUPPORTED;
        }
    std::vector<b200_op> list;
    if (ctx->opt_fusion) fuse(ops, n_ops, list);
    else list.assign(ops, ops + n_ops);

    if (!ctx->opt_cuda_graphs || n_ops < 8) return run_list(ctx, list);

    // ---- CUDA graph replay keyed on the exact op list ----
    if (!ctx->graph_cache) ctx->graph_cache = new GraphCache();
    GraphCache &gc = *ctx->graph_cache;
    for (auto &e : gc.entries) {
        if (ops_equal(e.ops, ops, n_ops)) {
            if (e.exec) {
                CUDA_TRY(cudaGraphLaunch(e.exec, ctx->stream));
                ctx->launches += 1;
                e.hits++;
                return B200_OK;
            }
            // second sighting: capture now.  Scratch must already be large enough (first run grew it).
            cudaGraph_t g = nullptr;
            ctx->capturing = true;
            CUDA_TRY(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            int rc = run_list(ctx, list);
            cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
            ctx->capturing = false;
            if (rc || ce != cudaSuccess || !g) {
                cudaGetLastError();
                if (g) cudaGraphDestroy(g);
                return run_list(ctx, list);          // fall back to eager launches
            }
            if (cudaGraphInstantiate(&e.exec, g, 0) != cudaSuccess) { cudaGetLastError(); e.exec = nullptr; cudaGraphDestroy(g); return run_list(ctx, list); }
            cudaGraphDestroy(g);
            CUDA_TRY(cudaGraphLaunch(e.exec, ctx->stream));
            ctx->launches += 1;
            return B200_OK;
        }
    }
    if (gc.entries.size() >= 64) {           // bounded cache: drop the oldest
        if (gc.entries.front().exec) cudaGraphExecDestroy(gc.entries.front().exec);
        gc.entries.erase(gc.entries.begin());
    }
    GraphCacheEntry ne;
    ne.ops.assign(ops, ops + n_ops);
    gc.entries.push_back(std::move(ne));
    return run_list(ctx, list);
}
