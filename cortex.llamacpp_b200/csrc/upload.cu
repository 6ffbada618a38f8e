// upload.cu -- GGUF weights -> device: pipelined, sliced upload from pageable (mmap'd) host memory.
//
// Replaces the upload loop of llama_model_loader::load_all_data (llama-model-loader.cpp:895-1071: 4 pinned 1 MiB staging buffers +
// events, one tensor after the other through ggml_backend_tensor_set_async) for the case that decides time-to-first-token on a
// multi-GPU B200 box: each tensor-parallel rank takes only ITS slice of every tensor out of the mmap'd file
//   * a row range       (wq/wk/wv/gate/up and the output matrix are split by rows), or
//   * a byte range of every row  (wo/down are split along K: whole quant blocks, so the slice of a row is contiguous bytes)
// and packs it into a dense device tensor.  Pageable memory cannot be DMA'd directly: a few host threads gather the slice into a
// ring of pinned chunks while the copy engine drains the previous chunks (cudaMemcpyAsync + one event per chunk), so the file is
// read once, at min(page-cache memcpy, PCIe) speed, with no intermediate full-size host copy.  The bytes on the device are the
// GGUF bytes (get_tensor reads them back unchanged).
#include "common.cuh"
#include <thread>
#include <vector>

namespace {
constexpr size_t UP_CHUNK = (size_t)16 << 20;
constexpr int UP_RING = 4;
constexpr int UP_THREADS = 4;

struct UploadRing {
    uint8_t *buf[UP_RING] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[UP_RING];
    bool used[UP_RING] = {false, false, false, false};
    int device = -1;
};
thread_local UploadRing g_ring;

int ring_init(int device) {
    if (g_ring.device == device) return B200_OK;
    if (g_ring.device >= 0) {
        for (int i = 0; i < UP_RING; i++) { cudaFreeHost(g_ring.buf[i]); cudaEventDestroy(g_ring.ev[i]); g_ring.used[i] = false; }
        g_ring.device = -1;
    }
    for (int i = 0; i < UP_RING; i++) {
        CUDA_TRY(cudaMallocHost((void **)&g_ring.buf[i], UP_CHUNK));
        CUDA_TRY(cudaEventCreateWithFlags(&g_ring.ev[i], cudaEventDisableTiming));
    }
    g_ring.device = device;
    return B200_OK;
}

// gather bytes [off, off + n) of the packed slice (rows of `col_bytes` taken `src_row_stride` apart) into dst
void gather(uint8_t *dst, const uint8_t *src, size_t src_row_stride, size_t col_bytes, size_t off, size_t n) {
    while (n) {
        const size_t r = off / col_bytes, c = off % col_bytes;
        const size_t take = std::min(n, col_bytes - c);
        memcpy(dst, src + r * src_row_stride + c, take);
        dst += take; off += take; n -= take;
    }
}
}  // namespace

// dst (device, dense [n_rows][col_bytes]) <- src[r * src_row_stride + col_off .. + col_bytes) for r in [row0, row0 + n_rows)
extern "C" int b200_upload_slice(b200_ctx *ctx, void *dst, const void *src, size_t src_row_stride, int64_t row0, int64_t n_rows, size_t col_off,
                                 size_t col_bytes) {
    if (!ctx || !dst || !src) return B200_ERR_FAILED;
    if (n_rows <= 0 || col_bytes == 0) return B200_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ring_init(ctx->device);
    if (rc) return rc;
    const uint8_t *base = (const uint8_t *)src + (size_t)row0 * src_row_stride + col_off;
    const size_t total = (size_t)n_rows * col_bytes;
    int slot = 0;
    for (size_t off = 0; off < total; off += UP_CHUNK, slot = (slot + 1) % UP_RING) {
        const size_t n = std::min(UP_CHUNK, total - off);
        if (g_ring.used[slot]) CUDA_TRY(cudaEventSynchronize(g_ring.ev[slot]));       // the copy engine is done with this chunk
        uint8_t *stage = g_ring.buf[slot];
        if (n >= ((size_t)1 << 20)) {
            std::thread th[UP_THREADS];
            const size_t per = (n + UP_THREADS - 1) / UP_THREADS;
            for (int t = 0; t < UP_THREADS; t++) {
                const size_t b = std::min(n, t * per), e = std::min(n, (t + 1) * per);
                th[t] = std::thread([=] { if (e > b) gather(stage + b, base, src_row_stride, col_bytes, off + b, e - b); });
            }
            for (int t = 0; t < UP_THREADS; t++) th[t].join();
        } else {
            gather(stage, base, src_row_stride, col_bytes, off, n);
        }
        CUDA_TRY(cudaMemcpyAsync((uint8_t *)dst + off, stage, n, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaEventRecord(g_ring.ev[slot], ctx->stream));
        g_ring.used[slot] = true;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < UP_RING; i++) g_ring.used[i] = false;
    return B200_OK;
}

// whole tensor (or a byte range of it)
extern "C" int b200_upload(b200_ctx *ctx, void *dst, const void *src, size_t size) {
    return b200_upload_slice(ctx, dst, src, size, 0, 1, 0, size);
}
