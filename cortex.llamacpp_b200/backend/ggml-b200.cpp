// ggml-b200.cpp -- the ggml backend ("B200") that llama.cpp / cortex.llamacpp loads.
//
// This file is the reference-side half of the drop-in boundary: it implements ggml's five backend vtables
// (llama.cpp/ggml/src/ggml-backend-impl.h:17-207) -- reg, device, buffer type, buffer, backend(stream) -- and the
// two dlopen entry points ggml_backend_load() looks for (ggml-backend-reg.cpp:227-271), and forwards everything to
// the C ABI in include/ggml_b200.h.  It contains no kernels and no CUDA calls: graph_compute() translates the
// ggml_cgraph into a flat b200_op list and hands it to b200_graph_compute(), which owns fusion, CUDA graphs and all
// device code.  LlamaEngine / LlamaServerContext / llama.cpp stay unchanged.
//
// Role-for-role it replaces ggml-cuda.cu:514-1112 (buffers), :2338-2431 (backend), :2790-2813 (events),
// :2815-3500 (device + reg), without split buffers (tensor parallelism lives below the C ABI).
#include "ggml.h"
#include "ggml-backend.h"
#include "ggml-backend-impl.h"
#include "ggml-impl.h"

#include "ggml_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#define B200_MAX_DEVICES 16

#define B200_EXPORT __attribute__((visibility("default")))
extern "C" B200_EXPORT ggml_backend_reg_t ggml_backend_b200_reg(void);

namespace {

struct device_ctx {
    int device;
    std::string name, description;
};
struct buft_ctx {
    int device;
    std::string name;
};
struct buffer_ctx {
    int device;
    void *base;
};
struct backend_ctx {
    int device;
    std::string name;
    b200_ctx *ctx;
    std::vector<b200_op> ops;     // reused translation buffer
};

ggml_guid_t backend_guid() {
    static ggml_guid guid = {0xb2, 0x00, 0x10, 0x0a, 0x5c, 0x47, 0x4e, 0x21, 0x9d, 0x3e, 0x11, 0x6f, 0x70, 0x42, 0xb2, 0x00};
    return &guid;
}

// ------------------------------------------------------------------------------------------------ translation
bool buffer_is_split(ggml_backend_buffer_t buffer);
void to_b200_tensor(const ggml_tensor *t, b200_tensor &o) {
    memset(&o, 0, sizeof(o));
    if (!t) return;
    o.data = t->data;
    o.type = (int32_t)t->type;
    o.flags = (t->buffer && t->buffer->usage == GGML_BACKEND_BUFFER_USAGE_WEIGHTS) ? B200_TENSOR_FLAG_WEIGHT : 0u;
    if (t->buffer && buffer_is_split(t->buffer)) {          // row-split weight: the b200_split made by split_buffer_init_tensor
        o.data = t->extra;
        o.flags |= B200_TENSOR_FLAG_SPLIT | B200_TENSOR_FLAG_WEIGHT;
    }
    for (int i = 0; i < 4; i++) { o.ne[i] = t->ne[i]; o.nb[i] = t->nb[i]; }
}

bool is_view_op(enum ggml_op op) {
    return op == GGML_OP_NONE || op == GGML_OP_RESHAPE || op == GGML_OP_VIEW || op == GGML_OP_PERMUTE || op == GGML_OP_TRANSPOSE;
}

// returns false if the node has no B200 equivalent
bool translate(const ggml_tensor *node, b200_op &o) {
    memset(&o, 0, sizeof(o));
    int op = B200_OP_NONE;
    int nsrc = 0;
    switch (node->op) {
        case GGML_OP_MUL_MAT: op = B200_OP_MUL_MAT; nsrc = 2; break;
        case GGML_OP_MUL_MAT_ID: op = B200_OP_MUL_MAT_ID; nsrc = 3; break;
        case GGML_OP_FLASH_ATTN_EXT: op = B200_OP_FLASH_ATTN_EXT; nsrc = 4; break;
        case GGML_OP_RMS_NORM: op = B200_OP_RMS_NORM; nsrc = 1; break;
        case GGML_OP_ROPE: op = B200_OP_ROPE; nsrc = 3; break;
        case GGML_OP_CPY: case GGML_OP_DUP: op = B200_OP_CPY; nsrc = 1; break;
        case GGML_OP_CONT: op = B200_OP_CONT; nsrc = 1; break;
        case GGML_OP_ADD: op = B200_OP_ADD; nsrc = 2; break;
        case GGML_OP_SUB: op = B200_OP_SUB; nsrc = 2; break;
        case GGML_OP_MUL: op = B200_OP_MUL; nsrc = 2; break;
        case GGML_OP_DIV: op = B200_OP_DIV; nsrc = 2; break;
        case GGML_OP_GET_ROWS: op = B200_OP_GET_ROWS; nsrc = 2; break;
        case GGML_OP_SOFT_MAX: op = B200_OP_SOFT_MAX; nsrc = 2; break;
        case GGML_OP_ARGSORT: op = B200_OP_ARGSORT; nsrc = 1; break;
        case GGML_OP_SUM_ROWS: op = B200_OP_SUM_ROWS; nsrc = 1; break;
        case GGML_OP_ARGMAX: op = B200_OP_ARGMAX; nsrc = 1; break;
        case GGML_OP_SCALE: op = B200_OP_SCALE; nsrc = 1; break;
        case GGML_OP_UNARY:
            nsrc = 1;
            switch (ggml_get_unary_op(node)) {
                case GGML_UNARY_OP_SILU: op = B200_OP_SILU; break;
                case GGML_UNARY_OP_GELU: op = B200_OP_GELU; break;
                case GGML_UNARY_OP_RELU: op = B200_OP_RELU; break;
                case GGML_UNARY_OP_TANH: op = B200_OP_TANH; break;
                case GGML_UNARY_OP_SIGMOID: op = B200_OP_SIGMOID; break;
                default: return false;
            }
            break;
        default: return false;
    }
    o.op = op;
    o.n_src = nsrc;
    static_assert(sizeof(node->op_params) >= sizeof(o.params), "op_params too small");
    memcpy(o.params, node->op_params, sizeof(o.params));
    to_b200_tensor(node, o.dst);
    for (int i = 0; i < nsrc && i < B200_MAX_SRC; i++) to_b200_tensor(node->src[i], o.src[i]);
    return true;
}

// ------------------------------------------------------------------------------------------------ buffer
void buffer_free(ggml_backend_buffer_t buffer) {
    buffer_ctx *c = (buffer_ctx *)buffer->context;
    b200_free(c->device, c->base);
    delete c;
}
void *buffer_get_base(ggml_backend_buffer_t buffer) { return ((buffer_ctx *)buffer->context)->base; }

enum ggml_status buffer_init_tensor(ggml_backend_buffer_t buffer, ggml_tensor *tensor) {
    buffer_ctx *c = (buffer_ctx *)buffer->context;
    if (tensor->view_src != NULL) return GGML_STATUS_SUCCESS;
    if (ggml_is_quantized(tensor->type) && buffer->usage != GGML_BACKEND_BUFFER_USAGE_COMPUTE) {
        // zero the tail padding the streaming kernels may read (never used in arithmetic, but keep it defined)
        const size_t orig = ggml_nbytes(tensor);
        const size_t padded = ggml_backend_buft_get_alloc_size(buffer->buft, tensor);
        if (padded > orig) b200_memset(c->device, (char *)tensor->data + orig, 0, padded - orig);
    }
    return GGML_STATUS_SUCCESS;
}
void buffer_memset_tensor(ggml_backend_buffer_t buffer, ggml_tensor *tensor, uint8_t value, size_t offset, size_t size) {
    b200_memset(((buffer_ctx *)buffer->context)->device, (char *)tensor->data + offset, value, size);
}
// per-device context used by the buffer interface for the pipelined weight upload (buffers are not tied to a backend stream)
b200_ctx *upload_ctx(int device) {
    static std::mutex mu;
    static b200_ctx *ctxs[16] = {nullptr};
    std::lock_guard<std::mutex> lk(mu);
    if (device < 0 || device >= 16) return nullptr;
    if (!ctxs[device]) ctxs[device] = b200_ctx_create(device);
    return ctxs[device];
}
void buffer_set_tensor(ggml_backend_buffer_t buffer, ggml_tensor *tensor, const void *data, size_t offset, size_t size) {
    const int device = ((buffer_ctx *)buffer->context)->device;
    // model load: `data` points into the mmap'd GGUF; big tensors go through the pinned-chunk pipeline (upload.cu) instead of one
    // synchronous pageable cudaMemcpy (llama-model-loader.cpp:1019-1060 does the same with 4 x 1 MiB staging buffers)
    if (size >= ((size_t)8 << 20)) {
        b200_ctx *uc = upload_ctx(device);
        if (uc && b200_upload(uc, (char *)tensor->data + offset, data, size) == B200_OK) return;
    }
    if (b200_memcpy_h2d(device, (char *)tensor->data + offset, data, size) != B200_OK)
        GGML_ABORT("ggml-b200: set_tensor failed: %s", b200_last_error());
}
void buffer_get_tensor(ggml_backend_buffer_t buffer, const ggml_tensor *tensor, void *data, size_t offset, size_t size) {
    if (b200_memcpy_d2h(((buffer_ctx *)buffer->context)->device, data, (const char *)tensor->data + offset, size) != B200_OK)
        GGML_ABORT("ggml-b200: get_tensor failed: %s", b200_last_error());
}
bool buffer_is_b200(ggml_backend_buffer_t buffer);
bool buffer_cpy_tensor(ggml_backend_buffer_t buffer, const ggml_tensor *src, ggml_tensor *dst) {
    if (src->buffer && buffer_is_b200(src->buffer) && ggml_is_contiguous(src) && ggml_is_contiguous(dst)) {
        buffer_ctx *s = (buffer_ctx *)src->buffer->context, *d = (buffer_ctx *)buffer->context;
        return b200_memcpy_d2d(d->device, dst->data, s->device, src->data, ggml_nbytes(src)) == B200_OK;
    }
    return false;
}
void buffer_clear(ggml_backend_buffer_t buffer, uint8_t value) {
    buffer_ctx *c = (buffer_ctx *)buffer->context;
    b200_memset(c->device, c->base, value, buffer->size);
}
const ggml_backend_buffer_i buffer_iface = {
    /* .free_buffer   = */ buffer_free,
    /* .get_base      = */ buffer_get_base,
    /* .init_tensor   = */ buffer_init_tensor,
    /* .memset_tensor = */ buffer_memset_tensor,
    /* .set_tensor    = */ buffer_set_tensor,
    /* .get_tensor    = */ buffer_get_tensor,
    /* .cpy_tensor    = */ buffer_cpy_tensor,
    /* .clear         = */ buffer_clear,
    /* .reset         = */ NULL,
};
bool buffer_is_b200(ggml_backend_buffer_t buffer) { return buffer->iface.free_buffer == buffer_free; }

// ------------------------------------------------------------------------------------------------ buffer type
const char *buft_get_name(ggml_backend_buffer_type_t buft) { return ((buft_ctx *)buft->context)->name.c_str(); }
ggml_backend_buffer_t buft_alloc_buffer(ggml_backend_buffer_type_t buft, size_t size) {
    buft_ctx *c = (buft_ctx *)buft->context;
    void *p = b200_malloc(c->device, size);
    if (!p) {
        GGML_LOG_ERROR("%s: allocating %.2f MiB on device %d failed: %s\n", __func__, size / 1024.0 / 1024.0, c->device, b200_last_error());
        return NULL;          // caller handles (ggml-cuda.cu:654-660)
    }
    return ggml_backend_buffer_init(buft, buffer_iface, new buffer_ctx{c->device, p}, size);
}
size_t buft_get_alignment(ggml_backend_buffer_type_t) { return 128; }
size_t buft_get_alloc_size(ggml_backend_buffer_type_t, const ggml_tensor *tensor) {
    return b200_alloc_size((int32_t)tensor->type, tensor->ne, ggml_nbytes(tensor));
}
bool buft_is_host(ggml_backend_buffer_type_t) { return false; }
const ggml_backend_buffer_type_i buft_iface = {
    buft_get_name, buft_alloc_buffer, buft_get_alignment, /* get_max_size */ NULL, buft_get_alloc_size, buft_is_host,
};
bool buft_is_b200(ggml_backend_buffer_type_t buft) { return buft->iface.get_name == buft_get_name; }

ggml_backend_buffer_type_t device_buffer_type(int device) {
    static std::mutex mu;
    static ggml_backend_buffer_type bufts[B200_MAX_DEVICES];
    static bool init = false;
    std::lock_guard<std::mutex> lock(mu);
    if (!init) {
        for (int i = 0; i < B200_MAX_DEVICES; i++) {
            bufts[i].iface = buft_iface;
            bufts[i].device = i < (int)ggml_backend_reg_dev_count(ggml_backend_b200_reg()) ? ggml_backend_reg_dev_get(ggml_backend_b200_reg(), i) : NULL;
            bufts[i].context = new buft_ctx{i, "B200" + std::to_string(i)};
        }
        init = true;
    }
    return &bufts[device];
}

// ------------------------------------------------------------------------------------------------ split buffer type (tensor parallelism)
// ggml_backend_split_buffer_type (reference: ggml-cuda.cu:723-1050), the interface llama.cpp uses for --split-mode row
// (llama-model.cpp:326-355): a weight matrix allocated in this buffer type is cut by ROWS over all B200 devices.  init_tensor allocates
// one dense shard per device, set_tensor uploads every device's row range straight from the mmap'd file (b200_upload_slice), and
// MUL_MAT on such a weight runs on all GPUs at once below the C ABI (B200_TENSOR_FLAG_SPLIT, mulmat.cu op_mul_mat_split).
struct split_buft_ctx {
    int main_device;
    float split[B200_MAX_DEVICES + 1];        // cumulative fractions: device d owns [split[d], split[d+1]) of the rows
    std::string name;
};
struct split_extra { b200_split sp; };
struct split_buffer_ctx {
    std::vector<split_extra *> extras;
    ~split_buffer_ctx() {
        for (split_extra *e : extras) {
            for (int i = 0; i < e->sp.n_dev; i++) if (e->sp.shard[i]) b200_free(e->sp.device[i], e->sp.shard[i]);
            delete e;
        }
    }
};
constexpr int64_t SPLIT_ROUNDING = 128;       // shard boundaries fall on multiples of the GEMM row tile
void split_rows(const split_buft_ctx *c, const ggml_tensor *t, int n_dev, int64_t *row_low /* [n_dev + 1] */) {
    const int64_t nrows = ggml_nrows(t);
    row_low[0] = 0;
    for (int d = 1; d < n_dev; d++) {
        int64_t r = (int64_t)(nrows * c->split[d]);
        r -= r % SPLIT_ROUNDING;
        row_low[d] = r < row_low[d - 1] ? row_low[d - 1] : r;
    }
    row_low[n_dev] = nrows;
}
int split_n_dev() { return (int)ggml_backend_reg_dev_count(ggml_backend_b200_reg()); }

const char *split_buft_get_name(ggml_backend_buffer_type_t buft) { return ((split_buft_ctx *)buft->context)->name.c_str(); }
bool buft_is_split(ggml_backend_buffer_type_t buft) { return buft->iface.get_name == split_buft_get_name; }
void split_buffer_free(ggml_backend_buffer_t buffer) { delete (split_buffer_ctx *)buffer->context; }
void *split_buffer_get_base(ggml_backend_buffer_t) { return (void *)0x1000; }      // never dereferenced: the shards live in the tensor extras
enum ggml_status split_buffer_init_tensor(ggml_backend_buffer_t buffer, ggml_tensor *tensor) {
    GGML_ASSERT(tensor->view_src == nullptr);                  // views of split tensors are not supported (as in the reference)
    split_buffer_ctx *bc = (split_buffer_ctx *)buffer->context;
    const split_buft_ctx *c = (const split_buft_ctx *)buffer->buft->context;
    split_extra *e = new split_extra();
    memset(&e->sp, 0, sizeof(e->sp));
    const int n = split_n_dev();
    e->sp.n_dev = n;
    split_rows(c, tensor, n, e->sp.row_low);
    const size_t rb = ggml_row_size(tensor->type, tensor->ne[0]);
    for (int d = 0; d < n; d++) {
        e->sp.device[d] = d;
        const int64_t nr = e->sp.row_low[d + 1] - e->sp.row_low[d];
        if (nr <= 0) continue;
        e->sp.shard[d] = b200_malloc(d, (size_t)nr * rb + 512);           // + slack: kernels read whole 16-byte lines
        if (!e->sp.shard[d]) { GGML_LOG_ERROR("ggml-b200: split shard of %s on device %d: %s\n", tensor->name, d, b200_last_error()); delete e; return GGML_STATUS_ALLOC_FAILED; }
        b200_memset(d, (char *)e->sp.shard[d] + (size_t)nr * rb, 0, 512);
    }
    bc->extras.push_back(e);
    tensor->extra = &e->sp;
    return GGML_STATUS_SUCCESS;
}
b200_ctx *upload_ctx(int device);
void split_buffer_set_tensor(ggml_backend_buffer_t, ggml_tensor *tensor, const void *data, size_t offset, size_t size) {
    GGML_ASSERT(offset == 0 && size == ggml_nbytes(tensor));   // split tensors are set in their entirety (as in the reference)
    const b200_split *sp = (const b200_split *)tensor->extra;
    const size_t rb = ggml_row_size(tensor->type, tensor->ne[0]);
    for (int d = 0; d < sp->n_dev; d++) {
        const int64_t nr = sp->row_low[d + 1] - sp->row_low[d];
        if (nr <= 0) continue;
        b200_ctx *uc = upload_ctx(sp->device[d]);
        if (!uc || b200_upload_slice(uc, sp->shard[d], data, rb, sp->row_low[d], nr, 0, rb) != B200_OK)
            GGML_ABORT("ggml-b200: split set_tensor(%s) on device %d: %s", tensor->name, sp->device[d], b200_last_error());
    }
}
void split_buffer_get_tensor(ggml_backend_buffer_t, const ggml_tensor *tensor, void *data, size_t offset, size_t size) {
    GGML_ASSERT(offset == 0 && size == ggml_nbytes(tensor));
    const b200_split *sp = (const b200_split *)tensor->extra;
    const size_t rb = ggml_row_size(tensor->type, tensor->ne[0]);
    for (int d = 0; d < sp->n_dev; d++) {
        const int64_t nr = sp->row_low[d + 1] - sp->row_low[d];
        if (nr > 0 && b200_memcpy_d2h(sp->device[d], (char *)data + sp->row_low[d] * rb, sp->shard[d], (size_t)nr * rb) != B200_OK)
            GGML_ABORT("ggml-b200: split get_tensor: %s", b200_last_error());
    }
}
void split_buffer_clear(ggml_backend_buffer_t, uint8_t) {}
const ggml_backend_buffer_i split_buffer_iface = {
    /* .free_buffer   = */ split_buffer_free,
    /* .get_base      = */ split_buffer_get_base,
    /* .init_tensor   = */ split_buffer_init_tensor,
    /* .memset_tensor = */ NULL,
    /* .set_tensor    = */ split_buffer_set_tensor,
    /* .get_tensor    = */ split_buffer_get_tensor,
    /* .cpy_tensor    = */ NULL,
    /* .clear         = */ split_buffer_clear,
    /* .reset         = */ NULL,
};
// by buffer TYPE: llama.cpp probes supports_op with a zero-sized dummy buffer whose interface is empty (llama-model.cpp:232-236)
bool buffer_is_split(ggml_backend_buffer_t buffer) { return buffer->buft && buft_is_split(buffer->buft); }
size_t buft_get_alignment(ggml_backend_buffer_type_t);
bool buft_is_host(ggml_backend_buffer_type_t);
ggml_backend_buffer_t split_buft_alloc_buffer(ggml_backend_buffer_type_t buft, size_t size) {
    // the shards are allocated per tensor in init_tensor (their sizes depend on the rounded row split)
    return ggml_backend_buffer_init(buft, split_buffer_iface, new split_buffer_ctx(), size);
}
size_t split_buft_get_alloc_size(ggml_backend_buffer_type_t, const ggml_tensor *tensor) { return ggml_nbytes(tensor) + 512 * (size_t)split_n_dev(); }
const ggml_backend_buffer_type_i split_buft_iface = {
    split_buft_get_name, split_buft_alloc_buffer, buft_get_alignment, /* get_max_size */ NULL, split_buft_get_alloc_size, buft_is_host,
};
ggml_backend_buffer_type_t split_buffer_type(int main_device, const float *tensor_split) {
    static std::mutex mu;
    static std::vector<ggml_backend_buffer_type *> made;
    std::lock_guard<std::mutex> lock(mu);
    const int n = split_n_dev();
    if (main_device < 0 || main_device >= n) return NULL;
    split_buft_ctx *c = new split_buft_ctx();
    c->main_device = main_device;
    float sum = 0.0f;
    bool all_zero = true;
    for (int d = 0; d < n && tensor_split; d++) all_zero = all_zero && tensor_split[d] == 0.0f;
    for (int d = 0; d < n; d++) { c->split[d] = sum; sum += (tensor_split && !all_zero) ? tensor_split[d] : 1.0f; }
    for (int d = 0; d < n; d++) c->split[d] /= sum;
    c->split[n] = 1.0f;
    c->name = "B200" + std::to_string(main_device) + "_Split";
    for (ggml_backend_buffer_type *b : made) {
        const split_buft_ctx *o = (const split_buft_ctx *)b->context;
        if (o->main_device == main_device && memcmp(o->split, c->split, sizeof(float) * (n + 1)) == 0) { delete c; return b; }
    }
    ggml_backend_buffer_type *b = new ggml_backend_buffer_type{split_buft_iface, ggml_backend_reg_dev_get(ggml_backend_b200_reg(), main_device), c};
    made.push_back(b);
    return b;
}

// pinned host buffer type (CPU-side weights, the scheduler's CPU compute buffer)
const char *host_buft_name(ggml_backend_buffer_type_t) { return "B200_Host"; }
void host_buffer_free(ggml_backend_buffer_t buffer) { b200_host_free(buffer->context); }
ggml_backend_buffer_t host_buft_alloc(ggml_backend_buffer_type_t buft, size_t size) {
    void *p = b200_host_malloc(size);
    if (!p) return ggml_backend_buft_alloc_buffer(ggml_backend_cpu_buffer_type(), size);   // fall back to pageable host memory
    ggml_backend_buffer_t buffer = ggml_backend_cpu_buffer_from_ptr(p, size);
    buffer->buft = buft;
    buffer->iface.free_buffer = host_buffer_free;
    return buffer;
}
ggml_backend_buffer_type_t host_buffer_type() {
    static ggml_backend_buffer_type t = {
        {host_buft_name, host_buft_alloc, ggml_backend_cpu_buffer_type()->iface.get_alignment, NULL,
         ggml_backend_cpu_buffer_type()->iface.get_alloc_size, ggml_backend_cpu_buffer_type()->iface.is_host},
        ggml_backend_reg_dev_get(ggml_backend_b200_reg(), 0),
        nullptr,
    };
    return &t;
}

// ------------------------------------------------------------------------------------------------ backend (stream)
const char *backend_get_name(ggml_backend_t backend) { return ((backend_ctx *)backend->context)->name.c_str(); }
void backend_free(ggml_backend_t backend) {
    backend_ctx *c = (backend_ctx *)backend->context;
    b200_ctx_destroy(c->ctx);
    delete c;
    delete backend;
}
void backend_set_tensor_async(ggml_backend_t backend, ggml_tensor *tensor, const void *data, size_t offset, size_t size) {
    backend_ctx *c = (backend_ctx *)backend->context;
    if (b200_memcpy_h2d_async(c->ctx, (char *)tensor->data + offset, data, size) != B200_OK) GGML_ABORT("ggml-b200: %s", b200_last_error());
}
void backend_get_tensor_async(ggml_backend_t backend, const ggml_tensor *tensor, void *data, size_t offset, size_t size) {
    backend_ctx *c = (backend_ctx *)backend->context;
    if (b200_memcpy_d2h_async(c->ctx, data, (const char *)tensor->data + offset, size) != B200_OK) GGML_ABORT("ggml-b200: %s", b200_last_error());
}
void backend_synchronize(ggml_backend_t backend) {
    if (b200_synchronize(((backend_ctx *)backend->context)->ctx) != B200_OK) GGML_ABORT("ggml-b200: synchronize: %s", b200_last_error());
}
enum ggml_status backend_graph_compute(ggml_backend_t backend, ggml_cgraph *cgraph) {
    backend_ctx *c = (backend_ctx *)backend->context;
    c->ops.clear();
    for (int i = 0; i < cgraph->n_nodes; i++) {
        ggml_tensor *node = cgraph->nodes[i];
        if (ggml_is_empty(node) || is_view_op(node->op)) continue;
        b200_op op;
        if (!translate(node, op)) {
            GGML_LOG_ERROR("ggml-b200: op %s reached graph_compute but is not supported\n", ggml_op_desc(node));
            return GGML_STATUS_FAILED;
        }
        c->ops.push_back(op);
    }
    const int rc = b200_graph_compute(c->ctx, c->ops.data(), (int)c->ops.size());
    if (rc == B200_OK) return GGML_STATUS_SUCCESS;
    GGML_LOG_ERROR("ggml-b200: graph_compute failed: %s\n", b200_last_error());
    return rc == B200_ERR_ALLOC ? GGML_STATUS_ALLOC_FAILED : GGML_STATUS_FAILED;
}
void backend_event_record(ggml_backend_t backend, ggml_backend_event_t event) {
    b200_event_record(((backend_ctx *)backend->context)->ctx, (b200_event *)event->context);
}
void backend_event_wait(ggml_backend_t backend, ggml_backend_event_t event) {
    b200_event_wait(((backend_ctx *)backend->context)->ctx, (b200_event *)event->context);
}
const ggml_backend_i backend_iface = {
    /* .get_name           = */ backend_get_name,
    /* .free               = */ backend_free,
    /* .set_tensor_async   = */ backend_set_tensor_async,
    /* .get_tensor_async   = */ backend_get_tensor_async,
    /* .cpy_tensor_async   = */ NULL,
    /* .synchronize        = */ backend_synchronize,
    /* .graph_plan_create  = */ NULL,
    /* .graph_plan_free    = */ NULL,
    /* .graph_plan_update  = */ NULL,
    /* .graph_plan_compute = */ NULL,
    /* .graph_compute      = */ backend_graph_compute,
    /* .event_record       = */ backend_event_record,
    /* .event_wait         = */ backend_event_wait,
};

// ------------------------------------------------------------------------------------------------ device
const char *dev_get_name(ggml_backend_dev_t dev) { return ((device_ctx *)dev->context)->name.c_str(); }
const char *dev_get_description(ggml_backend_dev_t dev) { return ((device_ctx *)dev->context)->description.c_str(); }
void dev_get_memory(ggml_backend_dev_t dev, size_t *free, size_t *total) {
    b200_device_info(((device_ctx *)dev->context)->device, NULL, 0, free, total, NULL, NULL, NULL);
}
enum ggml_backend_dev_type dev_get_type(ggml_backend_dev_t) { return GGML_BACKEND_DEVICE_TYPE_GPU; }
void dev_get_props(ggml_backend_dev_t dev, ggml_backend_dev_props *props) {
    props->name = dev_get_name(dev);
    props->description = dev_get_description(dev);
    props->type = dev_get_type(dev);
    dev_get_memory(dev, &props->memory_free, &props->memory_total);
    props->caps = {/* async */ true, /* host_buffer */ getenv("GGML_B200_NO_PINNED") == nullptr, /* buffer_from_host_ptr */ false, /* events */ true};
}
ggml_backend_t dev_init_backend(ggml_backend_dev_t dev, const char *) {
    device_ctx *d = (device_ctx *)dev->context;
    b200_ctx *ctx = b200_ctx_create(d->device);
    if (!ctx) {
        GGML_LOG_ERROR("ggml-b200: failed to create a context on device %d: %s\n", d->device, b200_last_error());
        return NULL;
    }
    // decode steps replay a captured CUDA graph (the KV-store destinations are patched through a device table, graph.cu) and
    // kernels use programmatic dependent launch; both can be switched off for debugging
    const char *eg = getenv("GGML_B200_GRAPHS"), *ep = getenv("GGML_B200_PDL");
    b200_set_option(ctx, "cuda_graphs", eg ? atoi(eg) : 1);
    b200_set_option(ctx, "pdl", ep ? atoi(ep) : 1);
    ggml_backend_t backend = new ggml_backend{backend_guid(), backend_iface, dev, new backend_ctx{d->device, d->name, ctx, {}}};
    return backend;
}
ggml_backend_buffer_type_t dev_get_buffer_type(ggml_backend_dev_t dev) { return device_buffer_type(((device_ctx *)dev->context)->device); }
ggml_backend_buffer_type_t dev_get_host_buffer_type(ggml_backend_dev_t) { return host_buffer_type(); }

bool dev_supports_op(ggml_backend_dev_t dev, const ggml_tensor *op) {
    device_ctx *d = (device_ctx *)dev->context;
    // every source that lives in one of our buffers must live on THIS device
    for (int i = 0; i < GGML_MAX_SRC; i++) {
        if (op->src[i] && op->src[i]->buffer && buft_is_b200(op->src[i]->buffer->buft) &&
            ((buft_ctx *)op->src[i]->buffer->buft->context)->device != d->device) return false;
    }
    // a row-split weight can only be the src0 of a MUL_MAT driven from its main device (as in the reference, ggml-cuda.cu:3080-3100)
    for (int i = 0; i < GGML_MAX_SRC; i++) {
        if (!op->src[i] || !op->src[i]->buffer || !buffer_is_split(op->src[i]->buffer)) continue;
        if (op->op != GGML_OP_MUL_MAT || i != 0 || ((split_buft_ctx *)op->src[i]->buffer->buft->context)->main_device != d->device) return false;
    }
    if (is_view_op(op->op)) return true;
    b200_op o;
    if (!translate(op, o)) return false;
    return b200_supports_op(d->device, &o) != 0;
}
bool dev_supports_buft(ggml_backend_dev_t dev, ggml_backend_buffer_type_t buft) {
    if (buft_is_split(buft)) return ((split_buft_ctx *)buft->context)->main_device == ((device_ctx *)dev->context)->device;
    return buft_is_b200(buft) && ((buft_ctx *)buft->context)->device == ((device_ctx *)dev->context)->device;
}
int64_t op_batch_size(const ggml_tensor *op) {
    switch (op->op) {
        case GGML_OP_GET_ROWS: return 0;
        case GGML_OP_MUL_MAT: return op->ne[1];
        case GGML_OP_MUL_MAT_ID: case GGML_OP_ROPE: return op->ne[2];
        default: return ggml_nrows(op);
    }
}
bool dev_offload_op(ggml_backend_dev_t, const ggml_tensor *op) {
    static const bool off = getenv("GGML_B200_NO_OFFLOAD") != nullptr;   // keep -ngl 0 runs purely on the CPU backend (parity baselines)
    return !off && op_batch_size(op) >= 32;
}   // as ggml-cuda.cu:3264-3285

ggml_backend_event_t dev_event_new(ggml_backend_dev_t dev) {
    b200_event *e = b200_event_create(((device_ctx *)dev->context)->device);
    if (!e) return NULL;
    return new ggml_backend_event{dev, e};
}
void dev_event_free(ggml_backend_dev_t, ggml_backend_event_t event) {
    b200_event_destroy((b200_event *)event->context);
    delete event;
}
void dev_event_synchronize(ggml_backend_dev_t, ggml_backend_event_t event) { b200_event_synchronize((b200_event *)event->context); }

const ggml_backend_device_i device_iface = {
    dev_get_name, dev_get_description, dev_get_memory, dev_get_type, dev_get_props, dev_init_backend, dev_get_buffer_type,
    dev_get_host_buffer_type, /* buffer_from_host_ptr */ NULL, dev_supports_op, dev_supports_buft, dev_offload_op,
    dev_event_new, dev_event_free, dev_event_synchronize,
};

// ------------------------------------------------------------------------------------------------ reg
struct reg_ctx {
    std::vector<ggml_backend_dev_t> devices;
};
const char *reg_get_name(ggml_backend_reg_t) { return "B200"; }
size_t reg_get_device_count(ggml_backend_reg_t reg) { return ((reg_ctx *)reg->context)->devices.size(); }
ggml_backend_dev_t reg_get_device(ggml_backend_reg_t reg, size_t index) {
    reg_ctx *c = (reg_ctx *)reg->context;
    GGML_ASSERT(index < c->devices.size());
    return c->devices[index];
}
ggml_backend_feature *get_features(ggml_backend_reg_t) {
    static ggml_backend_feature features[] = {{"ARCH", "sm_100a"}, {"ACT_QUANT", "cpu-exact(q8_0,q8_K)"}, {nullptr, nullptr}};
    return features;
}
void *reg_get_proc_address(ggml_backend_reg_t, const char *name) {
    if (strcmp(name, "ggml_backend_get_features") == 0) return (void *)get_features;
    // --split-mode row: llama.cpp asks the backend registry for this entry point (llama-model.cpp:326-355, ggml-cuda.cu:3412-3416)
    if (strcmp(name, "ggml_backend_split_buffer_type") == 0) return (void *)split_buffer_type;
    return NULL;
}
const ggml_backend_reg_i reg_iface = {reg_get_name, reg_get_device_count, reg_get_device, reg_get_proc_address};

}  // namespace

extern "C" {

ggml_backend_reg_t ggml_backend_b200_reg(void) {
    static ggml_backend_reg reg;
    static bool initialized = false;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!initialized) {
        reg_ctx *rc = new reg_ctx;
        const int n = b200_device_count();
        for (int i = 0; i < n && i < B200_MAX_DEVICES; i++) {
            char name[256] = "";
            b200_device_info(i, name, sizeof(name), NULL, NULL, NULL, NULL, NULL);
            device_ctx *dc = new device_ctx{i, "B200" + std::to_string(i), name};
            rc->devices.push_back(new ggml_backend_device{device_iface, &reg, dc});
        }
        reg = ggml_backend_reg{GGML_BACKEND_API_VERSION, reg_iface, rc};
        initialized = true;
    }
    return &reg;
}

// dlopen entry points (ggml_backend_load): score 0 => "not usable on this system"
B200_EXPORT int ggml_backend_score(void) { return b200_device_count() > 0 ? 100 : 0; }
B200_EXPORT ggml_backend_reg_t ggml_backend_init(void) { return ggml_backend_b200_reg(); }

// convenience for applications that link the backend directly (mirrors ggml_backend_cuda_init)
B200_EXPORT ggml_backend_t ggml_backend_b200_init(int device) {
    ggml_backend_reg_t reg = ggml_backend_b200_reg();
    if (device < 0 || (size_t)device >= ggml_backend_reg_dev_count(reg)) return NULL;
    return ggml_backend_dev_init(ggml_backend_reg_dev_get(reg, device), NULL);
}
B200_EXPORT bool ggml_backend_is_b200(ggml_backend_t backend) { return backend != NULL && ggml_guid_matches(backend->guid, backend_guid()); }

}  // extern "C"
