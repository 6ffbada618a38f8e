// dstep.h -- host interface of the persistent decode-step kernel (dstep.cu), used by graph.cu
#pragma once
#include "common.cuh"
#include "gemv.h"

enum { DS_GEMV = 0, DS_ATTN = 1, DS_COMBINE = 2, DS_COPY = 3 };

// one node of a decode step as graph.cu's matcher produced it (the same descriptors the per-launch kernels take)
struct DsNode {
    int kind = DS_GEMV;
    // DS_GEMV: a fused GEMV node (1-3 weight matrices sharing their activations, norm / swiglu prologue, residual epilogue)
    GemvSegDesc seg[GEMV_MAX_SEG]; int nseg = 0; int64_t K = 0; GemvActDesc act = {};
    // DS_ATTN: rope(q,k) + KV store + flash attention of ONE token (the EX_ROPE_STORE node and the FLASH_ATTN_EXT op after it)
    RopeStoreDesc rs = {}; b200_op fa = {};
    // DS_COPY: GET_ROWS of f32 rows (llama.cpp's inp_out_ids gather on the last layer)
    b200_op cp = {};
};

struct DsProgram;                     // device-resident program + launch geometry, cached per context

bool dstep_gemv_eligible(const b200_ctx *ctx, const GemvSegDesc *segs, int nseg, int64_t K, const GemvActDesc &act, int ncols);
bool dstep_attn_eligible(const b200_ctx *ctx, const RopeStoreDesc &rs, const b200_op &fa);
bool dstep_copy_eligible(const b200_op &op);
// builds (or finds in the context's cache) the device program for this node sequence; no launch.  Must not be called
// while the stream is capturing (it uploads the program).
int  dstep_prepare(b200_ctx *ctx, const std::vector<DsNode> &nodes, DsProgram **out);
int  dstep_launch(b200_ctx *ctx, DsProgram *prog);
void dstep_cache_free(b200_ctx *ctx);
