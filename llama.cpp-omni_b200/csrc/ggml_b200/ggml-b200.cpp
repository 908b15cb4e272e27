// ggml-b200.cpp — the drop-in boundary: a ggml backend plugin (libggml-b200.so) that the reference loads with
//   GGML_BACKEND_PATH=/path/libggml-b200.so        (ggml_backend_load_all, ggml/src/ggml-backend-reg.cpp:581-608)
// and that then sits behind the reference's own ggml_backend_t interface (ggml/src/ggml-backend-impl.h): llama.cpp-omni's graph
// executor (ggml_backend_sched), GGUF loader and omni pipeline stay untouched C++.  This file is HOST code only: every ggml op is
// mapped onto one entry point of the thin C-ABI of hand-written sm_100a kernels (include/b200_ops.h, libb200ops.so).
//
// Reference interface being filled in (what ggml-cuda.cu fills at :655-665, :724-731, :3191-3206, :3739-3755, :3847-3852):
//   ggml_backend_reg_i          ggml-backend-impl.h:194-204     -> b200_reg_*
//   ggml_backend_device_i       ggml-backend-impl.h:140-182     -> b200_dev_*
//   ggml_backend_buffer_type_i  ggml-backend-impl.h:17-29       -> b200_buft_*
//   ggml_backend_buffer_i       ggml-backend-impl.h:41-58       -> b200_buf_*
//   ggml_backend_i              ggml-backend-impl.h:87-120      -> b200_backend_*
// Entry points looked up with dlsym (ggml-backend-reg.cpp:257-273): ggml_backend_init, ggml_backend_score.
//
// Compiled against the reference's PUBLIC headers where they lie (-I$REF/ggml/include -I$REF/ggml/src); nothing is copied.
#include "ggml.h"
#include "ggml-backend.h"
#include "ggml-backend-impl.h"
#include "ggml-impl.h"                 // struct ggml_cgraph, ggml_node_has_n_uses (what the reference's own fusion checks use)
#include "../../../include/ggml-b200.h"
#include "../../../include/b200_ops.h"

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <chrono>
#include <string>
#include <unordered_map>
#include <vector>

#define B200_LOG(...) do { fprintf(stderr, "ggml-b200: " __VA_ARGS__); fputc('\n', stderr); } while (0)
// void interface functions cannot report errors: abort like the reference's CUDA_CHECK does (ggml-cuda/common.cuh)
#define CUDA_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
    fprintf(stderr, "ggml-b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorString(e_), __FILE__, __LINE__, #expr); abort(); } } while (0)

namespace {

constexpr int MAX_DEVICES = 16;

struct DeviceCtx {
    int index = 0;
    std::string name, description, pci_id;
    ggml_backend_buffer_type buft = {};
    ggml_backend_device dev = {};
};

struct BufferCtx {
    int device = 0;
    void * base = nullptr;
    void * staging = nullptr; size_t staging_size = 0;      // repack staging (weights only)
    std::mutex mu;
};

// 128-bit fingerprint of everything that decides whether a captured graph / a built engine can be replayed (see graph_key_of)
struct GraphKey {
    uint64_t h1 = 0, h2 = 0; int n = -1;
    bool operator==(const GraphKey & o) const { return h1 == o.h1 && h2 == o.h2 && n == o.n; }
    bool operator!=(const GraphKey & o) const { return !(*this == o); }
    void clear() { h1 = h2 = 0; n = -1; }
};

struct BackendCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    void * scratch = nullptr; size_t scratch_size = 0;      // activation records / split-KV partials
    std::string name;
    // decode-step CUDA graph (one per backend): re-captured whenever the node list or any tensor address / shape / parameter changes
    GraphKey graph_key, pending_key;
    cudaGraphExec_t graph_exec = nullptr;
    bool graphs_enabled = true;
    // whole-token decode engine (b200_decoder_*): built when a batch-1 graph matches the llama-family decoder pattern (match_decoder)
    void * engine = nullptr; int32_t engine_n_kv = 0;
    GraphKey engine_key, engine_reject_key;
    std::vector<int> engine_pre;                             // node INDICES still run per-op before the engine step (the KQ-mask cast); re-resolved from each graph
    cudaEvent_t hop_event = nullptr;                         // cpy_tensor_async: source-stream -> destination-stream ordering (re-recorded per hop)
    bool engine_enabled = true;
    // activation-tile reuse (B200_MM_REUSE_ACT): the tensor whose prepared activations the scratch currently holds, reset per graph_compute
    const ggml_tensor * scratch_act = nullptr; const void * scratch_act_data = nullptr; int scratch_act_type = -1;
    bool scratch_tiles_fused = false;                        // the tiles in the scratch came from a fused producer (fuse_tiles): the F32 tensor may not exist
    // q / k / v in one launch (b200_mul_mat_multi): the results of the MUL_MAT nodes computed EARLY (with the first of the group) wait here until the graph reaches
    // their node — their own dst may still hold a live tensor at that point (ggml-alloc reuses buffers in node order)
    void * hoist_buf = nullptr; size_t hoist_size = 0;
    const ggml_tensor * hoisted[2] = { nullptr, nullptr }; void * hoisted_at[2] = { nullptr, nullptr };
    long long host_ns = 0, host_min_ns = 0, host_calls = 0, host_nodes = 0, sync_ns = 0, sync_calls = 0, evsync_calls = 0;   // GGML_B200_HOST_TIMING
};

DeviceCtx g_devices[MAX_DEVICES];
int g_n_devices = -1;
ggml_backend_reg g_reg = {};

// ---------------------------------------------------------------------------------------------------------------- tensor views
bool type_known(enum ggml_type t) {
    switch (t) {
        case GGML_TYPE_F32: case GGML_TYPE_F16: case GGML_TYPE_BF16: case GGML_TYPE_Q4_0: case GGML_TYPE_Q8_0: case GGML_TYPE_Q4_K:
        case GGML_TYPE_Q5_K: case GGML_TYPE_Q6_K: case GGML_TYPE_I32: case GGML_TYPE_I64: return true;
        default: return false;
    }
}

// Weights whose native block is not a 16-byte multiple live in the PLANAR layout (include/b200_ops.h) when they sit, whole and
// 2-D, in a USAGE_WEIGHTS buffer: set_tensor scatters them, get_tensor gathers them back, MUL_MAT reads them planar.  The same
// predicate is evaluated on every path, so no per-tensor state is needed.  (Layout freedom: consumers only touch weight bytes
// through set/get_tensor — SURVEY.md §8b.)
bool is_planar(const ggml_tensor * t) {
    if (!t || !t->buffer || t->buffer->usage != GGML_BACKEND_BUFFER_USAGE_WEIGHTS) return false;
    if (!b200_repack_supported((int) t->type)) return false;
    return t->view_src == nullptr && t->ne[2] == 1 && t->ne[3] == 1 && ggml_is_contiguous(t);
}

b200_tensor view_of(const ggml_tensor * t) {
    b200_tensor v;
    v.data = t->data; v.type = (int32_t) t->type; v.layout = is_planar(t) ? B200_LAYOUT_PLANAR : B200_LAYOUT_NATIVE;
    for (int i = 0; i < 4; ++i) { v.ne[i] = t->ne[i]; v.nb[i] = (int64_t) t->nb[i]; }
    return v;
}

float fparam(const ggml_tensor * t, int i) { float f; memcpy(&f, (const int32_t *) t->op_params + i, sizeof(f)); return f; }
int32_t iparam(const ggml_tensor * t, int i) { return ((const int32_t *) t->op_params)[i]; }

// ---------------------------------------------------------------------------------------------------------------- buffers
void b200_buf_free(ggml_backend_buffer_t buffer) {
    BufferCtx * c = (BufferCtx *) buffer->context;
    cudaSetDevice(c->device);
    if (c->base) cudaFree(c->base);
    if (c->staging) cudaFree(c->staging);
    delete c;
}
void * b200_buf_get_base(ggml_backend_buffer_t buffer) { return ((BufferCtx *) buffer->context)->base; }

enum ggml_status b200_buf_init_tensor(ggml_backend_buffer_t buffer, ggml_tensor * tensor) {
    (void) buffer; (void) tensor;                                  // no per-tensor extras: layouts are derived (is_planar)
    return GGML_STATUS_SUCCESS;
}

void b200_buf_memset_tensor(ggml_backend_buffer_t buffer, ggml_tensor * tensor, uint8_t value, size_t offset, size_t size) {
    BufferCtx * c = (BufferCtx *) buffer->context;
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaMemset((char *) tensor->data + offset, value, size));
    CUDA_OK(cudaDeviceSynchronize());
}

void * staging_for(BufferCtx * c, size_t size) {
    if (c->staging_size < size) {
        if (c->staging) CUDA_OK(cudaFree(c->staging));
        c->staging_size = size < ((size_t) 4 << 20) ? ((size_t) 4 << 20) : size;
        CUDA_OK(cudaMalloc(&c->staging, c->staging_size));
    }
    return c->staging;
}

// uploads arrive in arbitrary (offset, size) chunks (llama-model-loader.cpp:1077-1093): planar weights are scattered chunk by chunk
void b200_buf_set_tensor(ggml_backend_buffer_t buffer, ggml_tensor * tensor, const void * data, size_t offset, size_t size) {
    BufferCtx * c = (BufferCtx *) buffer->context;
    CUDA_OK(cudaSetDevice(c->device));
    if (size == 0) return;
    if (is_planar(tensor)) {
        std::lock_guard<std::mutex> lock(c->mu);
        void * st = staging_for(c, size);
        CUDA_OK(cudaMemcpy(st, data, size, cudaMemcpyHostToDevice));
        const int64_t nblocks = ggml_nelements(tensor) / ggml_blck_size(tensor->type);
        const int rc = b200_repack_scatter((int) tensor->type, st, tensor->data, nblocks, (int64_t) offset, (int64_t) size, nullptr);
        if (rc) { B200_LOG("repack_scatter failed: %s", b200_error_string(rc)); abort(); }
        CUDA_OK(cudaStreamSynchronize(nullptr));
        return;
    }
    // the compute streams are cudaStreamNonBlocking (no implicit ordering with the legacy stream this copy runs on), and a copy from
    // pageable memory may return before the DMA has landed: finish it here (the reference: cudaMemcpyAsync + cudaStreamSynchronize(cudaStreamPerThread))
    CUDA_OK(cudaMemcpy((char *) tensor->data + offset, data, size, cudaMemcpyHostToDevice));
    CUDA_OK(cudaStreamSynchronize(nullptr));
}

void b200_buf_get_tensor(ggml_backend_buffer_t buffer, const ggml_tensor * tensor, void * data, size_t offset, size_t size) {
    BufferCtx * c = (BufferCtx *) buffer->context;
    CUDA_OK(cudaSetDevice(c->device));
    if (size == 0) return;
    if (is_planar(tensor)) {
        std::lock_guard<std::mutex> lock(c->mu);
        void * st = staging_for(c, size);
        const int64_t nblocks = ggml_nelements(tensor) / ggml_blck_size(tensor->type);
        const int rc = b200_repack_gather((int) tensor->type, tensor->data, st, nblocks, (int64_t) offset, (int64_t) size, nullptr);
        if (rc) { B200_LOG("repack_gather failed: %s", b200_error_string(rc)); abort(); }
        CUDA_OK(cudaMemcpy(data, st, size, cudaMemcpyDeviceToHost));
        return;
    }
    CUDA_OK(cudaMemcpy(data, (const char *) tensor->data + offset, size, cudaMemcpyDeviceToHost));
}

bool buffer_is_b200(ggml_backend_buffer_t buffer);

bool b200_buf_cpy_tensor(ggml_backend_buffer_t buffer, const ggml_tensor * src, ggml_tensor * dst) {
    (void) buffer;
    if (!src->buffer || !buffer_is_b200(src->buffer)) return false;                     // host sources go through set_tensor
    if (is_planar(src) != is_planar(dst) || ggml_nbytes(src) != ggml_nbytes(dst)) return false;
    BufferCtx * sc = (BufferCtx *) src->buffer->context, * dc = (BufferCtx *) dst->buffer->context;
    if (sc->device == dc->device) { CUDA_OK(cudaSetDevice(dc->device)); CUDA_OK(cudaMemcpy(dst->data, src->data, ggml_nbytes(src), cudaMemcpyDeviceToDevice)); }
    else CUDA_OK(cudaMemcpyPeer(dst->data, dc->device, src->data, sc->device, ggml_nbytes(src)));
    CUDA_OK(cudaStreamSynchronize(nullptr));                          // D2D copies are asynchronous to the host: the caller's next graph_compute runs on a non-blocking stream
    return true;
}

void b200_buf_clear(ggml_backend_buffer_t buffer, uint8_t value) {
    BufferCtx * c = (BufferCtx *) buffer->context;
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaMemset(c->base, value, buffer->size));
    CUDA_OK(cudaDeviceSynchronize());
}

const ggml_backend_buffer_i b200_buffer_iface = {
    /* free_buffer   */ b200_buf_free,
    /* get_base      */ b200_buf_get_base,
    /* init_tensor   */ b200_buf_init_tensor,
    /* memset_tensor */ b200_buf_memset_tensor,
    /* set_tensor    */ b200_buf_set_tensor,
    /* get_tensor    */ b200_buf_get_tensor,
    /* cpy_tensor    */ b200_buf_cpy_tensor,
    /* clear         */ b200_buf_clear,
    /* reset         */ nullptr,
};
bool buffer_is_b200(ggml_backend_buffer_t buffer) { return buffer->iface.free_buffer == b200_buf_free; }

// ---------------------------------------------------------------------------------------------------------------- buffer type
const char * b200_buft_get_name(ggml_backend_buffer_type_t buft) { return ((DeviceCtx *) buft->context)->name.c_str(); }

ggml_backend_buffer_t b200_buft_alloc(ggml_backend_buffer_type_t buft, size_t size) {
    DeviceCtx * d = (DeviceCtx *) buft->context;
    if (cudaSetDevice(d->index) != cudaSuccess) return nullptr;
    BufferCtx * c = new BufferCtx(); c->device = d->index;
    const size_t bytes = size ? size : 1;
    cudaError_t e = cudaMalloc(&c->base, bytes);
    if (e != cudaSuccess) {                                          // NULL on OOM: the allocator propagates it (ggml-cuda.cu:690-695)
        (void) cudaGetLastError();
        B200_LOG("allocating %.2f MiB on device %d failed: %s", size / 1024.0 / 1024.0, d->index, cudaGetErrorString(e));
        delete c; return nullptr;
    }
    return ggml_backend_buffer_init(buft, b200_buffer_iface, c, size);
}
size_t b200_buft_alignment(ggml_backend_buffer_type_t) { return 256; }
size_t b200_buft_max_size(ggml_backend_buffer_type_t) { return SIZE_MAX; }
size_t b200_buft_alloc_size(ggml_backend_buffer_type_t, const ggml_tensor * t) { return (ggml_nbytes(t) + 255) & ~(size_t) 255; }   // planar == native in bytes
bool   b200_buft_is_host(ggml_backend_buffer_type_t) { return false; }

// ---------------------------------------------------------------------------------------------------------------- pinned host buffer type
// What ggml_backend_cuda_host_buffer_type gives the reference (ggml-cuda.cu:1100-1156): llama allocates its output (logits) buffer and the CPU
// backend's compute buffer in the device's host buffer type, so that the per-token H2D input copies and the logits D2H are DMA transfers from
// page-locked memory.  The buffer behaves like a plain CPU buffer (ggml_backend_cpu_buffer_from_ptr); only allocation and free differ.
const char * b200_host_buft_get_name(ggml_backend_buffer_type_t) { return "B200_Host"; }
void b200_host_buf_free(ggml_backend_buffer_t buffer) { cudaFreeHost(buffer->context); }
ggml_backend_buffer_t b200_host_buft_alloc(ggml_backend_buffer_type_t buft, size_t size) {
    void * ptr = nullptr;
    if (getenv("GGML_B200_NO_PINNED") || cudaMallocHost(&ptr, size ? size : 1) != cudaSuccess) {
        (void) cudaGetLastError();
        return ggml_backend_buft_alloc_buffer(ggml_backend_cpu_buffer_type(), size);      // pageable fallback, still correct
    }
    ggml_backend_buffer_t buffer = ggml_backend_cpu_buffer_from_ptr(ptr, size);
    buffer->buft = buft;
    buffer->iface.free_buffer = b200_host_buf_free;
    return buffer;
}
size_t b200_host_buft_alignment(ggml_backend_buffer_type_t) { return 64; }
bool   b200_host_buft_is_host(ggml_backend_buffer_type_t) { return true; }
ggml_backend_buffer_type g_host_buft = {};
ggml_backend_buffer_type_t b200_dev_get_host_buffer_type(ggml_backend_dev_t dev) {
    static std::once_flag once;
    std::call_once(once, [dev] {
        g_host_buft.iface = { b200_host_buft_get_name, b200_host_buft_alloc, b200_host_buft_alignment, nullptr, nullptr, b200_host_buft_is_host };
        g_host_buft.device = dev; g_host_buft.context = nullptr;
    });
    return &g_host_buft;
}

// ---------------------------------------------------------------------------------------------------------------- supports_op
bool f32c(const ggml_tensor * t) { return t && t->type == GGML_TYPE_F32 && t->nb[0] == sizeof(float); }
bool floaty(const ggml_tensor * t) { return t && (t->type == GGML_TYPE_F32 || t->type == GGML_TYPE_F16 || t->type == GGML_TYPE_BF16); }
bool broadcastable(const ggml_tensor * a, const ggml_tensor * b) {
    for (int i = 0; i < 4; ++i) if (b->ne[i] == 0 || a->ne[i] % b->ne[i]) return false;
    return true;
}

int unary_map(enum ggml_unary_op u) {
    switch (u) {
        case GGML_UNARY_OP_SILU: return B200_SILU;       case GGML_UNARY_OP_GELU: return B200_GELU;         case GGML_UNARY_OP_RELU: return B200_RELU;
        case GGML_UNARY_OP_GELU_QUICK: return B200_GELU_QUICK; case GGML_UNARY_OP_TANH: return B200_TANH;   case GGML_UNARY_OP_SIGMOID: return B200_SIGMOID;
        case GGML_UNARY_OP_GELU_ERF: return B200_GELU_ERF; case GGML_UNARY_OP_NEG: return B200_NEG;         case GGML_UNARY_OP_EXP: return B200_EXP;
        case GGML_UNARY_OP_ABS: return B200_ABS;           case GGML_UNARY_OP_ELU: return B200_ELU;           case GGML_UNARY_OP_STEP: return B200_STEP;
        case GGML_UNARY_OP_SGN: return B200_SGN;           case GGML_UNARY_OP_HARDSWISH: return B200_HARDSWISH; case GGML_UNARY_OP_HARDSIGMOID: return B200_HARDSIGMOID;
        default: return -1;
    }
}

// Must depend on types / shapes / strides only: llama probes with dummy tensors whose buffer has size 0 (llama-model.cpp:286-291).
// debugging aid: GGML_B200_DISABLE_OPS=ROPE,FLASH_ATTN_EXT hands those ops back to the scheduler's CPU fallback (bisecting parity)
bool op_disabled(const ggml_tensor * op) {
    static const std::string list = [] { const char * e = getenv("GGML_B200_DISABLE_OPS"); return std::string(e ? e : ""); }();
    if (list.empty()) return false;
    const std::string name = ggml_op_name(op->op);
    size_t pos = 0;
    while (pos <= list.size()) {
        const size_t end = list.find(',', pos) == std::string::npos ? list.size() : list.find(',', pos);
        if (list.compare(pos, end - pos, name) == 0) return true;
        pos = end + 1;
    }
    return false;
}

bool b200_dev_supports_op(ggml_backend_dev_t, const ggml_tensor * op) {
    const ggml_tensor * s0 = op->src[0], * s1 = op->src[1];
    if (op_disabled(op)) return false;
    for (int i = 0; i < GGML_MAX_SRC; ++i) if (op->src[i] && !type_known(op->src[i]->type)) return false;
    if (!type_known(op->type)) return false;
    // a node without elements is never launched (is_noop): claim it whatever the op.  A prompt ubatch that returns no logits has an empty tail (GET_ROWS of zero output
    // rows, then ADD / norms / MUL_MATs over 0 columns); handing any of those to the CPU backend splits the graph there, and every CPU split behind a device split costs
    // a ggml_backend_synchronize of the device (no async copy to the CPU) — which serialised the scheduler's ubatch pipeline over the devices of `-sm layer`
    if (ggml_is_empty(op)) return true;
    switch (op->op) {
        case GGML_OP_NONE: case GGML_OP_RESHAPE: case GGML_OP_VIEW: case GGML_OP_PERMUTE: case GGML_OP_TRANSPOSE:
            return true;
        case GGML_OP_MUL_MAT: {
            if (!s0 || !s1 || ggml_is_transposed(s0) || ggml_is_transposed(s1)) return false;
            if (s0->view_src && b200_repack_supported((int) s0->type) && s0->buffer && s0->buffer->usage == GGML_BACKEND_BUFFER_USAGE_WEIGHTS) return false;   // a view into a planar weight
            b200_tensor w = view_of(s0), x = view_of(s1), d = view_of(op);
            return b200_mul_mat_supported(&w, &x, &d) != 0;
        }
        case GGML_OP_ADD: case GGML_OP_SUB: case GGML_OP_MUL: case GGML_OP_DIV:
            return floaty(s0) && floaty(s1) && floaty(op) && ggml_are_same_shape(s0, op) && broadcastable(s0, s1);
        case GGML_OP_RMS_NORM: case GGML_OP_NORM:
            return f32c(s0) && f32c(op) && ggml_are_same_shape(s0, op);
        case GGML_OP_IM2COL:
            return s1 && f32c(s1) && (op->type == GGML_TYPE_F16 || op->type == GGML_TYPE_F32) && ggml_is_contiguous(op);
        case GGML_OP_POOL_1D:
            return s0 && (s0->type == GGML_TYPE_F32 || s0->type == GGML_TYPE_F16) && op->type == GGML_TYPE_F32 && ggml_is_contiguous(s0) && ggml_is_contiguous(op) &&
                   iparam(op, 1) == iparam(op, 2) && iparam(op, 3) == 0 && (iparam(op, 0) == GGML_OP_POOL_MAX || iparam(op, 0) == GGML_OP_POOL_AVG);
        case GGML_OP_ROPE: {
            const int mode = iparam(op, 2), n_dims = iparam(op, 1);
            if (mode != 0 && mode != GGML_ROPE_TYPE_NEOX) return false;
            const bool f16io = s0 && s0->type == GGML_TYPE_F16 && op->type == GGML_TYPE_F16 && s0->nb[0] == 2 && op->nb[0] == 2;      // K-shift of the F16 cache
            if ((!f16io && (!f32c(s0) || !f32c(op))) || !s1 || s1->type != GGML_TYPE_I32 || !ggml_is_contiguous(s1)) return false;
            if (op->src[2] && (op->src[2]->type != GGML_TYPE_F32 || !ggml_is_contiguous(op->src[2]))) return false;
            return n_dims > 0 && n_dims % 2 == 0 && n_dims <= s0->ne[0] && s0->ne[0] % 2 == 0;
        }
        case GGML_OP_SET_ROWS:
            if (!s0 || !s1 || s0->type != GGML_TYPE_F32 || (s1->type != GGML_TYPE_I64 && s1->type != GGML_TYPE_I32)) return false;
            if (op->type == GGML_TYPE_Q8_0 || op->type == GGML_TYPE_Q4_0) { if (s0->ne[0] % 32 || s0->nb[0] != 4) return false; }          // quantised KV cache
            else if (!floaty(op)) return false;
            return s0->ne[0] == op->ne[0] && s0->ne[2] == op->ne[2] && s0->ne[3] == op->ne[3] && s1->ne[0] == s0->ne[1] &&
                   s1->ne[1] != 0 && s1->ne[2] != 0 && s0->ne[2] % s1->ne[1] == 0 && s0->ne[3] % s1->ne[2] == 0 && op->nb[0] == ggml_type_size(op->type);
        case GGML_OP_GET_ROWS:
            if (s0 && ggml_is_quantized(s0->type)) {                        // quantised token_embd: whole, contiguous tensors only (planar planes are addressed by block index)
                if (op->type != GGML_TYPE_F32 || !ggml_is_contiguous(s0) || s0->view_src) return false;
            } else if (!floaty(s0)) return false;
            return floaty(op) && s1 && s1->type == GGML_TYPE_I32 && s0->ne[0] == op->ne[0] && op->ne[1] == s1->ne[0] &&
                   op->ne[2] == s1->ne[1] && op->ne[3] == s1->ne[2];
        case GGML_OP_CPY: case GGML_OP_CONT: case GGML_OP_DUP: {
            if (!s0) return false;
            const bool ints = s0->type == GGML_TYPE_I32 && op->type == GGML_TYPE_I32;
            const bool to_q = s0->type == GGML_TYPE_F32 && (op->type == GGML_TYPE_Q8_0 || op->type == GGML_TYPE_Q4_0) && ggml_are_same_shape(s0, op) && s0->nb[0] == 4;
            const bool from_q = ggml_is_quantized(s0->type) && (op->type == GGML_TYPE_F32 || op->type == GGML_TYPE_F16) && ggml_are_same_shape(s0, op) && !is_planar(s0);
            if (to_q || from_q) return true;
            return (ints || (floaty(s0) && floaty(op))) && ggml_nelements(s0) == ggml_nelements(op);
        }
        case GGML_OP_SCALE:
            return floaty(s0) && floaty(op) && ggml_are_same_shape(s0, op);
        case GGML_OP_UNARY:
            return floaty(s0) && floaty(op) && ggml_are_same_shape(s0, op) && unary_map(ggml_get_unary_op(op)) >= 0;
        case GGML_OP_SQR: case GGML_OP_SQRT: case GGML_OP_SIN: case GGML_OP_COS: case GGML_OP_LOG: case GGML_OP_CLAMP: case GGML_OP_LEAKY_RELU:
            return floaty(s0) && floaty(op) && ggml_are_same_shape(s0, op);
        // ---- Token2Wav op set (csrc/ops_wave.cu): these graphs run without a scheduler, so anything missing here aborts the vocoder (token2wav-impl.cpp:1905-1916)
        case GGML_OP_CONCAT: {
            const int dim = iparam(op, 0);
            if (!s0 || !s1 || s0->type != s1->type || s0->type != op->type || !(floaty(s0) || s0->type == GGML_TYPE_I32) || dim < 0 || dim > 3) return false;
            for (int d = 0; d < 4; ++d) if (d == dim ? op->ne[d] != s0->ne[d] + s1->ne[d] : (s0->ne[d] != s1->ne[d] || op->ne[d] != s0->ne[d])) return false;
            return true;
        }
        case GGML_OP_REPEAT:
            return s0 && s0->type == op->type && (floaty(s0) || s0->type == GGML_TYPE_I32) && ggml_can_repeat(s0, op);
        case GGML_OP_ARANGE:
            return op->type == GGML_TYPE_F32 && ggml_is_contiguous(op) && ggml_nrows(op) == 1;
        case GGML_OP_SUM_ROWS:
            return f32c(s0) && f32c(op) && op->ne[0] == 1 && op->ne[1] == s0->ne[1] && op->ne[2] == s0->ne[2] && op->ne[3] == s0->ne[3];
        case GGML_OP_PAD: {
            if (!s0 || s0->type != GGML_TYPE_F32 || op->type != GGML_TYPE_F32) return false;
            for (int d = 0; d < 4; ++d) if (iparam(op, 2 * d) < 0 || iparam(op, 2 * d + 1) < 0 || op->ne[d] != s0->ne[d] + iparam(op, 2 * d) + iparam(op, 2 * d + 1)) return false;
            return true;
        }
        case GGML_OP_PAD_REFLECT_1D:
            return s0 && s0->type == GGML_TYPE_F32 && op->type == GGML_TYPE_F32 && iparam(op, 0) >= 0 && iparam(op, 1) >= 0 && iparam(op, 0) < s0->ne[0] && iparam(op, 1) < s0->ne[0];
        case GGML_OP_CONV_TRANSPOSE_1D:
            return s0 && s1 && (s0->type == GGML_TYPE_F32 || s0->type == GGML_TYPE_F16) && s1->type == GGML_TYPE_F32 && op->type == GGML_TYPE_F32 && s0->ne[3] == 1 &&
                   s1->ne[2] * s1->ne[3] == 1 && s0->ne[2] == s1->ne[1] && iparam(op, 0) > 0 && iparam(op, 1) == 0 && iparam(op, 2) == 1;
        case GGML_OP_GLU: {
            const enum ggml_glu_op g = ggml_get_glu_op(op);
            if (g == GGML_GLU_OP_SWIGLU_OAI || !floaty(s0) || !floaty(op)) return false;
            if (s1) return floaty(s1) && ggml_are_same_shape(s0, s1) && ggml_are_same_shape(s1, op);
            return s0->ne[0] == 2 * op->ne[0] && s0->ne[1] == op->ne[1] && s0->ne[2] == op->ne[2] && s0->ne[3] == op->ne[3];
        }
        case GGML_OP_SOFT_MAX: {
            if (!f32c(s0) || !f32c(op) || op->src[2] || fparam(op, 1) != 0.0f || s0->ne[0] > 24576) return false;
            if (s1 && ((s1->type != GGML_TYPE_F32 && s1->type != GGML_TYPE_F16) || s1->ne[0] != s0->ne[0] || s1->ne[1] < s0->ne[1] ||
                       s1->ne[2] == 0 || s1->ne[3] == 0 || s0->ne[2] % s1->ne[2] || s0->ne[3] % s1->ne[3] || s1->nb[0] != ggml_type_size(s1->type))) return false;
            return true;
        }
        case GGML_OP_FLASH_ATTN_EXT: {
            if (op->src[4] || fparam(op, 1) != 0.0f || fparam(op, 2) != 0.0f) return false;          // sinks / ALiBi / softcap: not on this path
            if (!s0 || !s1 || !op->src[2] || s0->ne[1] * s0->ne[3] > 65535) return false;
            b200_tensor q = view_of(s0), k = view_of(s1), v = view_of(op->src[2]), d = view_of(op), m;
            if (op->src[3]) m = view_of(op->src[3]);
            return b200_flash_attn_supported(&q, &k, &v, op->src[3] ? &m : nullptr, &d) != 0;
        }
        default:
            return false;
    }
}

// ---------------------------------------------------------------------------------------------------------------- graph compute
void * scratch_for(BackendCtx * c, size_t bytes) {
    if (bytes <= c->scratch_size) return c->scratch;
    CUDA_OK(cudaStreamSynchronize(c->stream));
    // the captured decode graph has the old scratch address baked into its kernel arguments: it must not be replayed after the free
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    c->graph_key.clear(); c->pending_key.clear();
    if (c->scratch) CUDA_OK(cudaFree(c->scratch));
    c->scratch_size = (bytes + ((size_t) 8 << 20)) & ~(((size_t) 1 << 20) - 1);
    CUDA_OK(cudaMalloc(&c->scratch, c->scratch_size));
    return c->scratch;
}

// scratch a node needs (so that it can be sized BEFORE a stream capture starts)
size_t node_scratch_bytes(const ggml_tensor * n) {
    if (n->op == GGML_OP_MUL_MAT) { b200_tensor w = view_of(n->src[0]), x = view_of(n->src[1]); return b200_mul_mat_scratch_bytes(&w, &x); }
    if (n->op == GGML_OP_FLASH_ATTN_EXT) { b200_tensor q = view_of(n->src[0]), k = view_of(n->src[1]); return b200_flash_attn_scratch_bytes(&q, &k); }
    return 0;
}

bool is_noop(const ggml_tensor * n) {
    return n->op == GGML_OP_NONE || n->op == GGML_OP_RESHAPE || n->op == GGML_OP_VIEW || n->op == GGML_OP_PERMUTE || n->op == GGML_OP_TRANSPOSE || ggml_is_empty(n);
}

// number of times `t` is read by nodes of the graph (views of t count as reads of t's consumers, conservatively as uses)
int n_uses(const ggml_cgraph * g, const ggml_tensor * t) {
    int n = 0;
    const int nn = ggml_graph_n_nodes((ggml_cgraph *) g);
    for (int i = 0; i < nn; ++i) {
        const ggml_tensor * c = ggml_graph_node((ggml_cgraph *) g, i);
        for (int s = 0; s < GGML_MAX_SRC; ++s) if (c->src[s] == t) ++n;
        if (c->view_src == t) ++n;
    }
    return n;
}

// One node -> one C-ABI call.  `next` (may be null) lets RMS_NORM absorb the MUL by the norm weight that follows it in every llama-family
// graph (the reference fuses the same pair, ggml-cuda.cu ggml_cuda_can_fuse / norm.cu:510-660).  Returns the number of nodes consumed.
// true when `n` (node i of g) is read exactly once in the WHOLE graph the scheduler split this view from: a graph view shares its parent's
// use counts (ggml_graph_view), which is what the reference's own fusion test reads (ggml_node_has_n_uses, ggml-impl.h:570); a count taken
// over the split alone would miss a consumer in another split / on another backend
bool single_use(const ggml_cgraph * g, int i, const ggml_tensor * n) {
    if (g->use_counts && g->visited_hash_set.keys) return ggml_node_has_n_uses(g, i, 1);
    return n_uses(g, n) == 1 && !n->view_src;
}

// does this MUL_MAT take the tensor-core path with prepared F16 activation tiles in the scratch?  (same routing conditions as mmq_tc_supported / mm_f16_tc_supported in
// csrc/mmq_tc.cu: native 16-byte-multiple quant blocks with back-to-back rows, planar planes, or F16 rows; more than 8 activation columns)
bool mm_tc_class(const ggml_tensor * n) {
    const ggml_tensor * s0 = n->src[0], * s1 = n->src[1];
    if (n->op != GGML_OP_MUL_MAT || !s0 || !s1 || s1->type != GGML_TYPE_F32) return false;
    const bool tc_native = (s0->type == GGML_TYPE_Q4_K || s0->type == GGML_TYPE_Q5_K) && s0->nb[1] == ggml_row_size(s0->type, s0->ne[0]);
    const bool tc_planar = (s0->type == GGML_TYPE_Q6_K || s0->type == GGML_TYPE_Q8_0 || s0->type == GGML_TYPE_Q4_0) && is_planar(s0);
    return (tc_native || tc_planar) && (uintptr_t) s0->data % 16 == 0 && s1->ne[1] > 8 && s1->ne[2] * s1->ne[3] == 1 && s0->ne[0] % 256 == 0;
}

// GGML_B200_NO_TILE_FUSION=1 switches the n-token fusions off (producer -> MUL_MAT tiles, MUL_MAT -> residual ADD epilogue); read per call: the parity harness
// toggles it between two runs of one process (llama_parity mode 8)
// (read ONCE per graph_compute into g_fusion_off: a getenv per node costs more host time than the launch it decides about)
thread_local bool g_fusion_off = false;
void fusion_refresh() { const char * env = getenv("GGML_B200_NO_TILE_FUSION"); g_fusion_off = env && atoi(env) != 0; }
inline bool fusion_off() { return g_fusion_off; }

bool is_weight(const ggml_tensor * t);
int run_node(BackendCtx * c, const ggml_cgraph * g, int node_idx, int & rc, bool allow_fuse = false) {
    ggml_tensor * n = ggml_graph_node((ggml_cgraph *) g, node_idx);
    ggml_tensor * next = allow_fuse && node_idx + 1 < ggml_graph_n_nodes((ggml_cgraph *) g) ? ggml_graph_node((ggml_cgraph *) g, node_idx + 1) : nullptr;
    void * st = c->stream;
    const ggml_tensor * s0 = n->src[0], * s1 = n->src[1];
    rc = B200_OK;
    switch (n->op) {
        case GGML_OP_MUL_MAT: {
            b200_tensor w = view_of(s0), x = view_of(s1), d = view_of(n);
            const size_t sb = b200_mul_mat_scratch_bytes(&w, &x);
            void * scratch_before = c->scratch;
            void * sc = sb ? scratch_for(c, sb) : nullptr;
            // q/k/v and gate/up share their input: the second and third projection reuse the F16 activation tiles the first one left in scratch
            // (only for the tensor-core path: > 8 columns of q4_K / planar q6_K; the preparation does not depend on which of the two types)
            // (same routing conditions as mmq_tc_supported in csrc/mmq_tc.cu: native 16-byte-multiple blocks with back-to-back rows, or planar planes)
            const bool tc_class = mm_tc_class(n);
            // activation class held in the scratch after this MUL_MAT: 1 = F16 tiles (tensor-core path), 2 / 3 = q8_K / q8_0 records of a matvec (<= 8 columns, no batch dims)
            const bool rec_class = !tc_class && ggml_is_quantized(s0->type) && s1->ne[1] <= 8 && s1->ne[2] * s1->ne[3] == 1 && s0->ne[2] * s0->ne[3] == 1;
            const bool kq = s0->type == GGML_TYPE_Q4_K || s0->type == GGML_TYPE_Q5_K || s0->type == GGML_TYPE_Q6_K;
            const int act_type = tc_class ? 1 : rec_class ? (kq ? 2 : 3) : 0;
            const bool reuse = act_type != 0 && !(act_type > 1 && fusion_off()) && sc == scratch_before && c->scratch_act == s1 && c->scratch_act_data == s1->data && c->scratch_act_type == act_type;
            c->scratch_act = act_type ? s1 : nullptr; c->scratch_act_data = s1->data; c->scratch_act_type = act_type;
            // ---- decode graphs (ONE activation column; the per-op route of everything the whole-token engine does not take: F16 / Q8_0 models, other head sizes,
            // a quantised KV cache): launches are what a token costs there, so adjacent nodes share one
            const bool one_col = allow_fuse && !fusion_off() && s1->ne[1] == 1 && s1->ne[2] * s1->ne[3] == 1 && s0->ne[2] * s0->ne[3] == 1 && is_weight(s0) &&
                                 (ggml_is_quantized(s0->type) || s0->type == GGML_TYPE_F16) && ggml_is_contiguous(n);
            if (one_col) {
                const int nn = ggml_graph_n_nodes((ggml_cgraph *) g);
                // [MUL_MAT gate|up, MUL_MAT up|gate, GLU(gate, up)] -> b200_mul_mat_glu
                ggml_tensor * n1 = node_idx + 1 < nn ? ggml_graph_node((ggml_cgraph *) g, node_idx + 1) : nullptr, * n2 = node_idx + 2 < nn ? ggml_graph_node((ggml_cgraph *) g, node_idx + 2) : nullptr;
                if (n1 && n2 && n1->op == GGML_OP_MUL_MAT && n1->src[1] == s1 && is_weight(n1->src[0]) && n2->op == GGML_OP_GLU && n2->src[1] && iparam(n2, 1) == 0 &&
                    ((n2->src[0] == n && n2->src[1] == n1) || (n2->src[0] == n1 && n2->src[1] == n)) && !(n->flags & GGML_TENSOR_FLAG_OUTPUT) && !(n1->flags & GGML_TENSOR_FLAG_OUTPUT) &&
                    single_use(g, node_idx, n) && single_use(g, node_idx + 1, n1) && ggml_is_contiguous(n2) && n2->type == GGML_TYPE_F32) {
                    const ggml_tensor * gate = n2->src[0], * up = n2->src[1];
                    b200_tensor wg = view_of(gate->src[0]), wu = view_of(up->src[0]), dg = view_of(n2);
                    rc = b200_mul_mat_glu((int) ggml_get_glu_op(n2), &wg, &wu, &x, &dg, sc, sb, reuse ? B200_MM_REUSE_ACT : 0, st);
                    if (rc != B200_ERR_UNSUPPORTED) return 3;                 // (the scratch holds x's q8 record either way: the bookkeeping above already says so)
                }
                // MUL_MAT -> ADD (the residual behind wo / ffn_down) -> the matvec's epilogue
                if (next && next->op == GGML_OP_ADD && (next->src[0] == n || next->src[1] == n) && next->src[0] != next->src[1] && !(n->flags & GGML_TENSOR_FLAG_OUTPUT) &&
                    single_use(g, node_idx, n)) {
                    const ggml_tensor * other = next->src[0] == n ? next->src[1] : next->src[0];
                    if (other->type == GGML_TYPE_F32 && next->type == GGML_TYPE_F32 && ggml_are_same_shape(other, n) && ggml_are_same_shape(next, n) && ggml_is_contiguous(other) && ggml_is_contiguous(next)) {
                        b200_tensor r = view_of(other), dn = view_of(next);
                        rc = b200_mul_mat_add(&w, &x, &r, &dn, sc, sb, reuse ? B200_MM_REUSE_ACT : 0, st);
                        if (rc != B200_ERR_UNSUPPORTED) return 2;
                    }
                }
            }
            // q / k / v: the later MUL_MATs over the same activations run with this one (one launch over the concatenated m-tiles); their results are parked in
            // hoist_buf and copied out when the graph reaches them (run_nodes)
            if (allow_fuse && tc_class && !fusion_off() && !c->hoisted[0] && x.ne[2] * x.ne[3] == 1) {
                const int nn = ggml_graph_n_nodes((ggml_cgraph *) g);
                const ggml_tensor * later[2] = { nullptr, nullptr }; int n_later = 0;
                bool direct[2] = { false, false };                               // the group member is the NEXT compute node: nothing runs in between, its own dst is safe to write now
                bool run = true;                                                 // every compute node since this one has been a member of the group
                for (int j = node_idx + 1; j < nn && j <= node_idx + 16 && n_later < 2; ++j) {
                    const ggml_tensor * t = ggml_graph_node((ggml_cgraph *) g, j);
                    if (is_noop(t)) continue;
                    if (t->op == GGML_OP_MUL_MAT && t->src[1] == s1 && t->src[0] != s0 && is_weight(t->src[0]) && mm_tc_class(t) && ggml_is_contiguous(t) && ggml_is_contiguous(n) &&
                        (t->src[0]->type == GGML_TYPE_Q4_K || t->src[0]->type == GGML_TYPE_Q5_K || t->src[0]->type == GGML_TYPE_Q6_K)) { direct[n_later] = run; later[n_later++] = t; }
                    else run = false;
                }
                if (n_later > 0 && is_weight(s0) && (s0->type == GGML_TYPE_Q4_K || s0->type == GGML_TYPE_Q5_K || s0->type == GGML_TYPE_Q6_K)) {
                    b200_tensor wv[3] = { w, view_of(later[0]->src[0]), n_later > 1 ? view_of(later[1]->src[0]) : w };
                    const b200_tensor * wp[3] = { &wv[0], &wv[1], &wv[2] };
                    if (b200_mul_mat_multi_merges(1 + n_later, wp, &x)) {
                        size_t need = 0, off[2] = { 0, 0 };
                        for (int i = 0; i < n_later; ++i) { off[i] = need; if (!direct[i]) need += (ggml_nbytes(later[i]) + 255) & ~(size_t) 255; }
                        if (need > c->hoist_size) {
                            CUDA_OK(cudaStreamSynchronize(c->stream));
                            if (c->hoist_buf) CUDA_OK(cudaFree(c->hoist_buf));
                            c->hoist_size = need; CUDA_OK(cudaMalloc(&c->hoist_buf, need));
                        }
                        b200_tensor dv[3] = { d, view_of(later[0]), n_later > 1 ? view_of(later[1]) : d };
                        size_t sbm = sb;
                        for (int i = 0; i < n_later; ++i) {
                            if (!direct[i]) dv[1 + i].data = (char *) c->hoist_buf + off[i];
                            const size_t sbi = b200_mul_mat_scratch_bytes(&wv[1 + i], &x); if (sbi > sbm) sbm = sbi;
                        }
                        if (sbm <= c->scratch_size || !reuse) {                  // growing the scratch would drop the prepared tiles
                            void * scm = scratch_for(c, sbm);
                            const b200_tensor * dp[3] = { &dv[0], &dv[1], &dv[2] };
                            rc = b200_mul_mat_multi(1 + n_later, wp, &x, dp, scm, sbm, reuse && scm == scratch_before ? B200_MM_REUSE_ACT : 0, st);
                            if (rc == B200_OK) for (int i = 0; i < n_later; ++i) { c->hoisted[i] = later[i]; c->hoisted_at[i] = direct[i] ? nullptr : dv[1 + i].data; }
                            return 1;
                        }
                    }
                }
            }
            // the residual ADD right behind wo / ffn_down rides in the GEMM epilogue (b200_mul_mat_add) when this MUL_MAT's result has no other reader
            if (tc_class && next && next->op == GGML_OP_ADD && (next->src[0] == n || next->src[1] == n) && next->src[0] != next->src[1] && !(n->flags & GGML_TENSOR_FLAG_OUTPUT) &&
                single_use(g, node_idx, n) && !fusion_off()) {
                const ggml_tensor * other = next->src[0] == n ? next->src[1] : next->src[0];
                if (other->type == GGML_TYPE_F32 && next->type == GGML_TYPE_F32 && ggml_are_same_shape(other, n) && ggml_are_same_shape(next, n) && ggml_is_contiguous(other) &&
                    ggml_is_contiguous(next) && ggml_is_contiguous(n)) {
                    b200_tensor r = view_of(other), dn = view_of(next);
                    rc = b200_mul_mat_add(&w, &x, &r, &dn, sc, sb, reuse ? B200_MM_REUSE_ACT : 0, st);
                    if (rc != B200_ERR_UNSUPPORTED) return 2;
                }
            }
            rc = b200_mul_mat_ex(&w, &x, &d, sc, sb, reuse ? B200_MM_REUSE_ACT : 0, st);
            return 1;
        }
        case GGML_OP_ADD: case GGML_OP_SUB: case GGML_OP_MUL: case GGML_OP_DIV: {
            b200_tensor a = view_of(s0), b = view_of(s1), d = view_of(n);
            const int op = n->op == GGML_OP_ADD ? B200_ADD : n->op == GGML_OP_SUB ? B200_SUB : n->op == GGML_OP_MUL ? B200_MUL : B200_DIV;
            rc = b200_binary(op, &a, &b, &d, st);
            return 1;
        }
        case GGML_OP_RMS_NORM: {
            b200_tensor x = view_of(s0), d = view_of(n);
            if (next && next->op == GGML_OP_MUL && (next->src[0] == n || next->src[1] == n) && !(n->flags & GGML_TENSOR_FLAG_OUTPUT) && single_use(g, node_idx, n)) {
                const ggml_tensor * wt = next->src[0] == n ? next->src[1] : next->src[0];
                // decode graphs: when a quantised matvec reads the normalised vector, the same launch also leaves its q8 record in the scratch (k_rms_norm_quantize),
                // and that MUL_MAT (and the ones after it on the same activations: q / k / v, gate / up) skip their quantisation pass
                if (!fusion_off() && f32c(wt) && wt->ne[0] == n->ne[0] && ggml_nelements(wt) == wt->ne[0] && n->ne[1] <= 8 && n->ne[2] * n->ne[3] == 1 && ggml_is_contiguous(next) &&
                    ggml_is_contiguous(wt)) {
                    const int nn = ggml_graph_n_nodes((ggml_cgraph *) g);
                    const ggml_tensor * mm = nullptr;
                    for (int j = node_idx + 2, seen = 0; j < nn && seen < 6 && !mm; ++j) {
                        const ggml_tensor * t = ggml_graph_node((ggml_cgraph *) g, j);
                        if (is_noop(t)) continue;
                        ++seen;
                        if (t->op == GGML_OP_MUL_MAT && t->src[1] == next && ggml_is_quantized(t->src[0]->type) && !mm_tc_class(t) && t->src[0]->ne[2] * t->src[0]->ne[3] == 1) mm = t;
                    }
                    if (mm) {
                        b200_tensor mw = view_of(mm->src[0]), mx = view_of(next);
                        const size_t sbm = b200_mul_mat_scratch_bytes(&mw, &mx);
                        void * scm = sbm ? scratch_for(c, sbm) : nullptr;
                        if (scm) {
                            rc = b200_rms_norm_quantize((const float *) s0->data, (int64_t) s0->nb[1] / 4, (const float *) wt->data, (float *) next->data, (int64_t) next->nb[1] / 4, scm,
                                                        (int) mm->src[0]->type, n->ne[0], n->ne[1], fparam(n, 0), st);
                            if (rc == B200_OK) {
                                const bool kq = mm->src[0]->type == GGML_TYPE_Q4_K || mm->src[0]->type == GGML_TYPE_Q5_K || mm->src[0]->type == GGML_TYPE_Q6_K;
                                c->scratch_act = next; c->scratch_act_data = next->data; c->scratch_act_type = kq ? 2 : 3;
                                return 2;
                            }
                            rc = B200_OK;
                        }
                    }
                }
                if (f32c(wt) && ggml_are_same_shape(next, n) && broadcastable(n, wt)) {
                    b200_tensor w = view_of(wt), dm = view_of(next);
                    rc = b200_rms_norm(&x, &w, nullptr, &dm, fparam(n, 0), st);
                    if (rc == B200_OK) return 2;
                }
            }
            rc = b200_rms_norm(&x, nullptr, nullptr, &d, fparam(n, 0), st);
            return 1;
        }
        case GGML_OP_NORM: { b200_tensor x = view_of(s0), d = view_of(n); rc = b200_norm(&x, &d, fparam(n, 0), st); return 1; }
        case GGML_OP_IM2COL: {
            b200_tensor k = view_of(s0), x = view_of(s1), d = view_of(n);
            rc = b200_im2col(&k, &x, &d, iparam(n, 0), iparam(n, 1), iparam(n, 2), iparam(n, 3), iparam(n, 4), iparam(n, 5), iparam(n, 6) == 1, st);
            return 1;
        }
        case GGML_OP_POOL_1D: { b200_tensor x = view_of(s0), d = view_of(n); rc = b200_pool_1d(&x, &d, iparam(n, 0), iparam(n, 1), iparam(n, 2), iparam(n, 3), st); return 1; }
        case GGML_OP_ROPE: {
            b200_tensor x = view_of(s0), d = view_of(n);
            b200_rope_params p;
            p.n_dims = iparam(n, 1); p.mode = iparam(n, 2); p.n_ctx_orig = iparam(n, 4);
            p.freq_base = fparam(n, 5); p.freq_scale = fparam(n, 6); p.ext_factor = fparam(n, 7); p.attn_factor = fparam(n, 8);
            p.beta_fast = fparam(n, 9); p.beta_slow = fparam(n, 10);
            rc = b200_rope(&x, (const int32_t *) s1->data, n->src[2] ? (const float *) n->src[2]->data : nullptr, &d, &p, st);
            return 1;
        }
        case GGML_OP_SET_ROWS: { b200_tensor a = view_of(s0), i = view_of(s1), d = view_of(n); rc = b200_set_rows(&a, &i, &d, st); return 1; }
        case GGML_OP_GET_ROWS: { b200_tensor a = view_of(s0), i = view_of(s1), d = view_of(n); rc = b200_get_rows(&a, &i, &d, st); return 1; }
        case GGML_OP_CPY: case GGML_OP_CONT: case GGML_OP_DUP: { b200_tensor a = view_of(s0), d = view_of(n); rc = b200_cpy(&a, &d, st); return 1; }
        case GGML_OP_SCALE: { b200_tensor a = view_of(s0), d = view_of(n); rc = b200_scale(&a, &d, fparam(n, 0), fparam(n, 1), st); return 1; }
        case GGML_OP_UNARY: { b200_tensor a = view_of(s0), d = view_of(n); rc = b200_unary_param(unary_map(ggml_get_unary_op(n)), &a, &d, 0.0f, 0.0f, st); return 1; }
        case GGML_OP_SQR:  { b200_tensor a = view_of(s0), d = view_of(n); rc = b200_unary(B200_SQR, &a, &d, st); return 1; }
        case GGML_OP_SQRT: { b200_tensor a = view_of(s0), d = view_of(n); rc = b200_unary(B200_SQRT, &a, &d, st); return 1; }
        case GGML_OP_SIN:  { b200_tensor a = view_of(s0), d = view_of(n); rc = b200_unary_param(B200_SIN, &a, &d, 0.0f, 0.0f, st); return 1; }
        case GGML_OP_COS:  { b200_tensor a = view_of(s0), d = view_of(n); rc = b200_unary_param(B200_COS, &a, &d, 0.0f, 0.0f, st); return 1; }
        case GGML_OP_LOG:  { b200_tensor a = view_of(s0), d = view_of(n); rc = b200_unary_param(B200_LOG, &a, &d, 0.0f, 0.0f, st); return 1; }
        case GGML_OP_CLAMP: { b200_tensor a = view_of(s0), d = view_of(n); rc = b200_unary_param(B200_CLAMP, &a, &d, fparam(n, 0), fparam(n, 1), st); return 1; }
        case GGML_OP_LEAKY_RELU: { b200_tensor a = view_of(s0), d = view_of(n); rc = b200_unary_param(B200_LEAKY_RELU, &a, &d, fparam(n, 0), 0.0f, st); return 1; }
        case GGML_OP_CONCAT: { b200_tensor a = view_of(s0), b = view_of(s1), d = view_of(n); rc = b200_concat(&a, &b, &d, iparam(n, 0), st); return 1; }
        case GGML_OP_REPEAT: { b200_tensor a = view_of(s0), d = view_of(n); rc = b200_repeat(&a, &d, st); return 1; }
        case GGML_OP_ARANGE: { b200_tensor d = view_of(n); rc = b200_arange(&d, fparam(n, 0), fparam(n, 2), st); return 1; }
        case GGML_OP_SUM_ROWS: { b200_tensor a = view_of(s0), d = view_of(n); rc = b200_sum_rows(&a, &d, st); return 1; }
        case GGML_OP_PAD: {
            b200_tensor a = view_of(s0), d = view_of(n);
            int32_t lr[8]; for (int i = 0; i < 8; ++i) lr[i] = iparam(n, i);
            rc = b200_pad(&a, &d, lr, st);
            return 1;
        }
        case GGML_OP_PAD_REFLECT_1D: { b200_tensor a = view_of(s0), d = view_of(n); rc = b200_pad_reflect_1d(&a, &d, iparam(n, 0), iparam(n, 1), st); return 1; }
        case GGML_OP_CONV_TRANSPOSE_1D: { b200_tensor k = view_of(s0), x = view_of(s1), d = view_of(n); rc = b200_conv_transpose_1d(&k, &x, &d, iparam(n, 0), st); return 1; }
        case GGML_OP_GLU: {
            b200_tensor a = view_of(s0), d = view_of(n), u;
            if (s1) u = view_of(s1);
            rc = b200_glu((int) ggml_get_glu_op(n), &a, s1 ? &u : nullptr, &d, iparam(n, 1), st);
            return 1;
        }
        case GGML_OP_SOFT_MAX: {
            b200_tensor a = view_of(s0), d = view_of(n), m;
            if (s1) m = view_of(s1);
            rc = b200_soft_max(&a, s1 ? &m : nullptr, &d, fparam(n, 0), fparam(n, 1), st);
            return 1;
        }
        case GGML_OP_FLASH_ATTN_EXT: {
            b200_tensor q = view_of(s0), k = view_of(s1), v = view_of(n->src[2]), d = view_of(n), m;
            if (n->src[3]) m = view_of(n->src[3]);
            const size_t sb = b200_flash_attn_scratch_bytes(&q, &k);
            c->scratch_act = nullptr;                                    // the attention kernels use the same scratch
            rc = b200_flash_attn(&q, &k, &v, n->src[3] ? &m : nullptr, &d, fparam(n, 0), fparam(n, 1), fparam(n, 2), sb ? scratch_for(c, sb) : nullptr, sb, st);
            return 1;
        }
        default:
            rc = B200_ERR_UNSUPPORTED;
            return 1;
    }
}

bool is_view_op(const ggml_tensor * t);
// ---- producer -> MUL_MAT fusion for n-token graphs -------------------------------------------------------------------------------------------------------------
// RMS_NORM(+MUL), GLU and FLASH_ATTN_EXT whose result only feeds tensor-core MUL_MATs write those MUL_MATs' F16 activation tiles straight into the scratch
// (b200_*_tiles) instead of F32 + one conversion pass per MUL_MAT; the MUL_MATs then run with B200_MM_REUSE_ACT through the scratch_act bookkeeping below.
// Whether the F32 tensor can be skipped is decided from the PARENT graph's use counts (a graph view shares them), so a reader in another split is never starved.
int graph_uses(const ggml_cgraph * g, const ggml_tensor * t) {
    if (!g->use_counts || !g->visited_hash_set.keys) return -1;
    const size_t pos = ggml_hash_find(&g->visited_hash_set, t);
    if (pos == GGML_HASHSET_FULL || !ggml_bitset_get(g->visited_hash_set.used, pos)) return -1;
    return g->use_counts[pos];
}

int fuse_tiles(BackendCtx * c, const ggml_cgraph * g, int i, int & rc) {
    const bool off = fusion_off();
    const int nn = ggml_graph_n_nodes((ggml_cgraph *) g);
    ggml_tensor * n = ggml_graph_node((ggml_cgraph *) g, i);
    rc = B200_OK;
    if (off) return 0;
    ggml_tensor * out = nullptr; int consumed = 0;
    if (n->op == GGML_OP_RMS_NORM && i + 1 < nn) {
        ggml_tensor * mul = ggml_graph_node((ggml_cgraph *) g, i + 1);
        if (mul->op != GGML_OP_MUL || (mul->src[0] != n && mul->src[1] != n) || (n->flags & GGML_TENSOR_FLAG_OUTPUT) || !single_use(g, i, n)) return 0;
        const ggml_tensor * wt = mul->src[0] == n ? mul->src[1] : mul->src[0];
        if (!f32c(wt) || !ggml_are_same_shape(mul, n) || wt->ne[0] != n->ne[0] || ggml_nelements(wt) != wt->ne[0]) return 0;
        out = mul; consumed = 2;
    } else if (n->op == GGML_OP_GLU && n->src[1]) { out = n; consumed = 1; }
    else if (n->op == GGML_OP_FLASH_ATTN_EXT) { out = n; consumed = 1; }
    else return 0;
    if (out->flags & GGML_TENSOR_FLAG_OUTPUT) return 0;
    // readers of `out` (through whole-tensor views) among the nodes that follow
    const ggml_tensor * alias[4] = { out, nullptr, nullptr, nullptr }; int n_alias = 1, seen[4] = { 0, 0, 0, 0 };
    const ggml_tensor * mm_first = nullptr; int n_mm = 0;
    size_t tile_bytes = 0;                                                  // the largest scratch any of the consumers will ask for: the scratch must not move under the tiles
    bool other_reader = false;
    for (int j = i + consumed; j < nn && j < i + consumed + 24; ++j) {
        const ggml_tensor * t = ggml_graph_node((ggml_cgraph *) g, j);
        int which = -1;
        for (int s = 0; s < GGML_MAX_SRC && which < 0; ++s) for (int a = 0; a < n_alias; ++a) if (t->src[s] && t->src[s] == alias[a]) { which = a; break; }
        if (which < 0) continue;
        if (is_view_op(t) && t->src[0] == alias[which] && ggml_nelements(t) == ggml_nelements(out) && t->data == out->data && n_alias < 4) { ++seen[which]; alias[n_alias++] = t; continue; }
        if (t->op == GGML_OP_MUL_MAT && t->src[1] == alias[which] && t->src[0] != alias[which] && mm_tc_class(t) && t->src[1]->data == out->data && ggml_is_contiguous(t->src[1]) &&
            (!mm_first || t->src[1] == mm_first->src[1])) {
            ++seen[which]; if (!mm_first) mm_first = t; ++n_mm;
            b200_tensor w = view_of(t->src[0]), xv = view_of(t->src[1]);
            const size_t sb = b200_mul_mat_scratch_bytes(&w, &xv);
            if (sb > tile_bytes) tile_bytes = sb;
            continue;
        }
        other_reader = true;
    }
    if (!mm_first) return 0;
    bool all_accounted = !other_reader;
    for (int a = 0; a < n_alias && all_accounted; ++a) all_accounted = graph_uses(g, alias[a]) == seen[a];
    const ggml_tensor * x2 = mm_first->src[1];                               // the [k, n] matrix the MUL_MATs read
    if (tile_bytes == 0) return 0;
    int rc2 = B200_ERR_UNSUPPORTED;
    if (n->op == GGML_OP_RMS_NORM) {
        ggml_tensor * mul = (ggml_tensor *) out;
        const ggml_tensor * wt = mul->src[0] == n ? mul->src[1] : mul->src[0];
        void * sc = scratch_for(c, tile_bytes);
        b200_tensor x = view_of(n->src[0]), wv = view_of(wt), d = view_of(mul);
        if (all_accounted) d.data = nullptr;                                 // nobody reads the F32 tensor
        rc2 = b200_rms_norm_tiles(&x, &wv, &d, sc, fparam(n, 0), c->stream);
    } else if (n->op == GGML_OP_GLU) {
        if (!all_accounted || iparam(n, 1) != 0) return 0;
        void * sc = scratch_for(c, tile_bytes);
        b200_tensor a = view_of(n->src[0]), u = view_of(n->src[1]);
        rc2 = b200_glu_tiles((int) ggml_get_glu_op(n), &a, &u, sc, c->stream);
    } else {
        if (!all_accounted || n->src[4] || fparam(n, 1) != 0.0f || fparam(n, 2) != 0.0f || !n->src[3]) return 0;
        b200_tensor q = view_of(n->src[0]), k = view_of(n->src[1]), v = view_of(n->src[2]), m = view_of(n->src[3]), d = view_of(n);
        const size_t fa_sb = b200_flash_attn_scratch_bytes(&q, &k), off2 = (tile_bytes + 255) & ~(size_t) 255;
        char * sc = (char *) scratch_for(c, off2 + fa_sb);
        rc2 = b200_flash_attn_tiles(&q, &k, &v, &m, &d, fparam(n, 0), fa_sb ? sc + off2 : nullptr, fa_sb, sc, c->stream);
    }
    if (rc2 == B200_ERR_UNSUPPORTED) return 0;                               // shapes the tile producers do not take: the plain path
    rc = rc2;
    if (rc == B200_OK) { c->scratch_act = x2; c->scratch_act_data = x2->data; c->scratch_act_type = 1; c->scratch_tiles_fused = true; }
    return consumed;
}

enum ggml_status run_nodes(BackendCtx * c, ggml_cgraph * g) {
    const int nn = ggml_graph_n_nodes(g);
    fusion_refresh();
    c->scratch_act = nullptr;
    c->hoisted[0] = c->hoisted[1] = nullptr;
    for (int i = 0; i < nn; ) {
        ggml_tensor * n = ggml_graph_node(g, i);
        if (is_noop(n)) { ++i; continue; }
        if (c->scratch_act && n->data == c->scratch_act_data) c->scratch_act = nullptr;     // an in-place op rewrites the tensor the tiles were made from
        if (n->op == GGML_OP_MUL_MAT && (n == c->hoisted[0] || n == c->hoisted[1])) {       // computed with the first MUL_MAT of its group: only the copy is left
            const int h = n == c->hoisted[0] ? 0 : 1;
            if (c->hoisted_at[h]) CUDA_OK(cudaMemcpyAsync(n->data, c->hoisted_at[h], ggml_nbytes(n), cudaMemcpyDeviceToDevice, c->stream));
            c->hoisted[h] = nullptr;
            ++i; continue;
        }
        int rc = 0;
        int used = (n->op == GGML_OP_RMS_NORM || n->op == GGML_OP_GLU || n->op == GGML_OP_FLASH_ATTN_EXT) && n->ne[1] * n->ne[2] >= 64 ? fuse_tiles(c, g, i, rc) : 0;
        if (used == 0) used = run_node(c, g, i, rc, true);               // RMS_NORM may absorb the MUL right behind it
        if (rc != B200_OK) {
            // graph_compute on an op supports_op rejected is a caller bug (the reference asserts, ggml-cuda.cu:3043-3047)
            B200_LOG("op %s (%s) failed: %s", ggml_op_name(n->op), n->name, b200_error_string(rc));
            return GGML_STATUS_FAILED;
        }
        i += used;
    }
    return GGML_STATUS_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------------- decode-engine matcher
// Recognise the batch-1 graph llm_build_qwen3 / llm_build_llama emit (src/llama-model.cpp:9287-9406, build_attn src/llama-graph.cpp:1546-1597,
// build_ffn :713-819, cpy_k / cpy_v src/llama-kv-cache.cpp:1021-1110) by DATAFLOW, walking back from the last node, and describe it as a
// b200_decode_desc: the whole split then runs as ONE persistent kernel (csrc/stream_decode.cu) instead of ~24 launches per layer.  Every
// compute node of the graph must be accounted for, otherwise the graph keeps the per-op path (LoRA, biases, control vectors, other archs).
bool is_view_op(const ggml_tensor * t) {
    return t->op == GGML_OP_RESHAPE || t->op == GGML_OP_VIEW || t->op == GGML_OP_PERMUTE || t->op == GGML_OP_TRANSPOSE;
}
const ggml_tensor * strip_views(const ggml_tensor * t) {       // the compute node behind a chain of whole-tensor views
    while (t && is_view_op(t)) { if (ggml_nelements(t) != ggml_nelements(t->src[0])) return nullptr; t = t->src[0]; }
    return t;
}
bool is_weight(const ggml_tensor * t) { return t && t->buffer && t->buffer->usage == GGML_BACKEND_BUFFER_USAGE_WEIGHTS && t->op == GGML_OP_NONE; }

struct DecoderMatch {
    std::vector<b200_decode_layer> layers;                    // filled back to front, reversed at the end
    b200_decode_desc d = {};
    std::vector<int> pre;                                      // node indices
    int n_matched = 0; int32_t n_kv = 0;
    float eps = -1.0f; bool have_rope = false;
    const ggml_tensor * pos = nullptr, * idx = nullptr, * mask = nullptr;
};

b200_weight weight_of(const ggml_tensor * w) { b200_weight r; r.data = w->data; r.type = (int32_t) w->type; r.layout = is_planar(w) ? B200_LAYOUT_PLANAR : B200_LAYOUT_NATIVE; return r; }

// y = MUL_MAT(w, x) with a 2-D weight and ONE activation column
bool m_mul_mat(const ggml_tensor * y, const ggml_tensor *& w, const ggml_tensor *& x, DecoderMatch & M) {
    if (!y || y->op != GGML_OP_MUL_MAT || !is_weight(y->src[0]) || y->src[0]->ne[2] != 1 || y->src[0]->ne[3] != 1 || ggml_nelements(y->src[1]) != y->src[0]->ne[0]) return false;
    if (y->flags & GGML_TENSOR_FLAG_OUTPUT) { /* only the logits may be an output; checked by the caller */ }
    w = y->src[0]; x = y->src[1]; ++M.n_matched;
    return true;
}
// y = MUL(RMS_NORM(x), w): build_norm (src/llama-graph.cpp:660-695)
bool m_norm(const ggml_tensor * y, const ggml_tensor *& x, const ggml_tensor *& w, DecoderMatch & M) {
    if (!y || y->op != GGML_OP_MUL) return false;
    const ggml_tensor * n = y->src[0], * wt = y->src[1];
    if (!n || n->op != GGML_OP_RMS_NORM || !is_weight(wt) || wt->type != GGML_TYPE_F32 || !ggml_is_contiguous(wt) || wt->ne[0] != n->ne[0] || ggml_nelements(wt) != wt->ne[0]) return false;
    const float eps = fparam(n, 0);
    if (M.eps >= 0.0f && eps != M.eps) return false;
    M.eps = eps; x = n->src[0]; w = wt; M.n_matched += 2;
    return true;
}
// t = ROPE([MUL(RMS_NORM(.), nw)] reshape(MUL_MAT(w, a)), pos): one of Qcur / Kcur
bool m_qk(const ggml_tensor * t, const ggml_tensor *& w, const ggml_tensor *& a, const ggml_tensor *& nw, DecoderMatch & M) {
    t = strip_views(t);
    if (!t || t->op != GGML_OP_ROPE || t->src[2] != nullptr || !t->src[1] || t->src[1]->type != GGML_TYPE_I32 || ggml_nelements(t->src[1]) != 1) return false;
    b200_rope_params p;
    p.n_dims = iparam(t, 1); p.mode = iparam(t, 2); p.n_ctx_orig = iparam(t, 4);
    p.freq_base = fparam(t, 5); p.freq_scale = fparam(t, 6); p.ext_factor = fparam(t, 7); p.attn_factor = fparam(t, 8); p.beta_fast = fparam(t, 9); p.beta_slow = fparam(t, 10);
    if (M.have_rope && memcmp(&p, &M.d.rope, sizeof(p)) != 0) return false;
    if (M.pos && M.pos != t->src[1]) return false;
    M.d.rope = p; M.have_rope = true; M.pos = t->src[1]; ++M.n_matched;
    const ggml_tensor * in = strip_views(t->src[0]);
    nw = nullptr;
    if (in && in->op == GGML_OP_MUL) { const ggml_tensor * x = nullptr; if (!m_norm(in, x, nw, M)) return false; in = strip_views(x); }
    return m_mul_mat(in, w, a, M);
}

bool match_decoder(const ggml_cgraph * g, DecoderMatch & M) {
    const int nn = ggml_graph_n_nodes((ggml_cgraph *) g);
    if (nn < 8) return false;
    int n_real = 0;
    std::unordered_map<const void *, const ggml_tensor *> set_rows;                 // cache tensor data -> SET_ROWS node writing it
    for (int i = 0; i < nn; ++i) {
        const ggml_tensor * n = ggml_graph_node((ggml_cgraph *) g, i);
        if (is_noop(n)) continue;
        ++n_real;
        if (n->op == GGML_OP_SET_ROWS) { if (!n->src[0] || !n->view_src) return false; set_rows[n->view_src->data] = n; }
    }
    const ggml_tensor * last = ggml_graph_node((ggml_cgraph *) g, nn - 1);
    const ggml_tensor * x = nullptr;                                                 // residual stream, walking backwards
    if (last->op == GGML_OP_MUL_MAT) {                                               // result_output = output . (RMS_NORM(x) * output_norm)
        const ggml_tensor * w = nullptr, * h = nullptr, * nw = nullptr;
        if (!m_mul_mat(last, w, h, M) || !m_norm(h, x, nw, M)) return false;
        M.d.lm_head = weight_of(w); M.d.out_norm = (const float *) nw->data; M.d.logits = (float *) last->data; M.d.hidden_out = (float *) h->data;
        M.d.n_vocab = (int32_t) w->ne[1];
    } else if (last->op == GGML_OP_ADD) { x = last; M.d.x_out = (float *) last->data; }
    else return false;

    while (x && x->op == GGML_OP_ADD) {
        b200_decode_layer L = {};
        // ---- l_out = ffn_out + ffn_inp;  ffn_out = down . swiglu(gate . f, up . f);  f = RMS_NORM(ffn_inp) * ffn_norm
        ++M.n_matched;
        const ggml_tensor * ffn_inp = x->src[1], * w = nullptr, * h = nullptr, * f = nullptr, * f2 = nullptr, * nw = nullptr, * nx = nullptr;
        if (!m_mul_mat(x->src[0], w, h, M)) return false;
        L.down = weight_of(w); M.d.n_ff = (int32_t) w->ne[0];
        if (!h || h->op != GGML_OP_GLU || ggml_get_glu_op(h) != GGML_GLU_OP_SWIGLU || !h->src[1] || iparam(h, 1) != 0) return false;
        ++M.n_matched;
        if (!m_mul_mat(h->src[0], w, f, M)) return false;
        L.gate = weight_of(w);
        if (!m_mul_mat(h->src[1], w, f2, M) || f2 != f) return false;
        L.up = weight_of(w);
        if (!m_norm(f, nx, nw, M) || nx != ffn_inp) return false;
        L.ffn_norm = (const float *) nw->data;
        // ---- ffn_inp = wo . attn + inpL   (last layer of a full model: both through an identity GET_ROWS of the single output row)
        if (!ffn_inp || ffn_inp->op != GGML_OP_ADD) return false;
        ++M.n_matched;
        const ggml_tensor * proj = ffn_inp->src[0], * inpL = ffn_inp->src[1];
        if (proj && proj->op == GGML_OP_GET_ROWS && inpL && inpL->op == GGML_OP_GET_ROWS) {
            if (proj->src[0]->ne[1] != 1 || inpL->src[0]->ne[1] != 1 || ggml_nelements(proj->src[1]) != 1 || proj->src[1] != inpL->src[1]) return false;
            proj = proj->src[0]; inpL = inpL->src[0]; M.n_matched += 2;
        }
        const ggml_tensor * kqv = nullptr;
        if (!m_mul_mat(proj, w, kqv, M)) return false;
        L.wo = weight_of(w); M.d.n_embd = (int32_t) w->ne[1];
        const ggml_tensor * fa = strip_views(kqv);
        if (!fa || fa->op != GGML_OP_FLASH_ATTN_EXT || fa->src[4] || fparam(fa, 1) != 0.0f || fparam(fa, 2) != 0.0f) return false;
        ++M.n_matched;
        const ggml_tensor * q = fa->src[0], * k = fa->src[1], * v = fa->src[2], * mask = fa->src[3];
        if (!q || !k || !v || !mask || mask->type != GGML_TYPE_F16 || k->type != GGML_TYPE_F16 || v->type != GGML_TYPE_F16) return false;
        if (q->ne[1] != 1 || q->ne[3] != 1 || k->ne[3] != 1) return false;
        const int64_t D = q->ne[0], H = q->ne[2], HK = k->ne[2];
        if (k->nb[2] != D * 2 || v->nb[2] != D * 2 || k->ne[1] != v->ne[1] || mask->ne[0] != k->ne[1]) return false;        // heads side by side inside a cache row
        if (M.mask && M.mask != mask) return false;
        if (M.n_kv && M.n_kv != (int32_t) k->ne[1]) return false;
        M.mask = mask; M.n_kv = (int32_t) k->ne[1];
        M.d.head_dim = (int32_t) D; M.d.n_head = (int32_t) H; M.d.n_head_kv = (int32_t) HK; M.d.attn_scale = fparam(fa, 0);
        L.k_cache = k->data; L.v_cache = v->data; L.k_row_bytes = (int64_t) k->nb[1]; L.v_row_bytes = (int64_t) v->nb[1];
        const ggml_tensor * a = nullptr, * a2 = nullptr, * qn = nullptr, * kn = nullptr;
        if (!m_qk(q, w, a, qn, M)) return false;
        L.wq = weight_of(w);
        // ---- KV-cache writes of this layer: SET_ROWS into the tensors the attention reads
        auto ik = set_rows.find(k->data), iv = set_rows.find(v->data);
        if (ik == set_rows.end() || iv == set_rows.end()) return false;
        const ggml_tensor * sk = ik->second, * sv = iv->second;
        if (sk->type != GGML_TYPE_F16 || sv->type != GGML_TYPE_F16 || (int64_t) sk->nb[1] != L.k_row_bytes || (int64_t) sv->nb[1] != L.v_row_bytes) return false;
        if (sk->src[1]->type != GGML_TYPE_I64 || ggml_nelements(sk->src[1]) != 1 || ggml_nelements(sv->src[1]) != 1 || sv->src[1]->type != GGML_TYPE_I64) return false;
        if (M.idx && M.idx != sk->src[1]) return false;
        M.idx = sk->src[1];                                                          // v_idxs holds the same cell index when V is not transposed
        M.n_matched += 2;
        if (!m_qk(sk->src[0], w, a2, kn, M) || a2 != a) return false;
        L.wk = weight_of(w);
        if (!m_mul_mat(strip_views(sv->src[0]), w, a2, M) || a2 != a) return false;
        L.wv = weight_of(w);
        if ((qn == nullptr) != (kn == nullptr)) return false;
        L.q_norm = qn ? (const float *) qn->data : nullptr; L.k_norm = kn ? (const float *) kn->data : nullptr;
        // ---- a = RMS_NORM(inpL) * attn_norm
        if (!m_norm(a, nx, nw, M) || nx != inpL) return false;
        L.attn_norm = (const float *) nw->data;
        M.layers.push_back(L);
        x = inpL;
    }
    if (!x || M.layers.empty() || !M.pos || !M.idx || !M.mask) return false;
    if (x->op != GGML_OP_NONE && !is_view_op(x)) return false;                      // the split's input: token embedding row or the previous stage's l_out
    if (x->type != GGML_TYPE_F32 || ggml_nelements(x) != M.d.n_embd || !ggml_is_contiguous(x)) return false;
    M.d.x_in = (const float *) x->data;
    if (M.mask->op == GGML_OP_CPY) {                                                 // the F32 -> F16 cast of the KQ mask stays a per-op launch
        int idx = -1;
        for (int i = 0; i < nn; ++i) if (ggml_graph_node((ggml_cgraph *) g, i) == M.mask) { idx = i; break; }
        if (idx < 0) return false;
        M.pre.push_back(idx); ++M.n_matched;
    }
    else if (M.mask->op != GGML_OP_NONE) return false;
    if (M.n_matched != n_real) return false;                                        // something in the graph is not part of the pattern
    for (int i = 0; i < nn; ++i) {                                                   // only the logits / result_norm may be graph outputs
        const ggml_tensor * n = ggml_graph_node((ggml_cgraph *) g, i);
        if ((n->flags & GGML_TENSOR_FLAG_OUTPUT) && n != last && (void *) n->data != (void *) M.d.hidden_out) return false;
    }
    std::vector<b200_decode_layer> fwd(M.layers.rbegin(), M.layers.rend());
    M.layers.swap(fwd);
    M.d.n_layer = (int32_t) M.layers.size(); M.d.layers = M.layers.data(); M.d.rms_eps = M.eps;
    M.d.pos = (const int32_t *) M.pos->data; M.d.kv_idx = (const int64_t *) M.idx->data; M.d.mask = M.mask->data;
    return true;
}

// The properties that decide whether a captured graph can be replayed (what the reference compares in
// ggml_cuda_graph_update_required / is_cuda_graph_update_required, ggml-cuda.cu:2800-2900): op, addresses, shapes, strides, op params —
// folded on the fly into two independent 64-bit multiply-xor hashes (no per-token allocation; ~900 nodes x ~50 words per decode graph).
// Four interleaved lanes per hash: one lane is a chain of dependent 64-bit multiplies (~5 cycles each), and ~50 k words per token made that chain ~70 us of the
// host's serial per-token time (GGML_B200_HOST_TIMING); word i goes to lane i & 3 (still position-sensitive), the lanes are folded at the end.
struct KeyHasher {
    uint64_t a[4] = { 0x9e3779b97f4a7c15ull, 0xbf58476d1ce4e5b9ull, 0x94d049bb133111ebull, 0x2545f4914f6cdd1dull };
    uint64_t b[4] = { 0xc2b2ae3d27d4eb4full, 0x165667b19e3779f9ull, 0xd6e8feb86659fd93ull, 0xff51afd7ed558ccdull };
    unsigned i = 0;
    inline void add(uint64_t v) {
        const unsigned l = i++ & 3;
        a[l] = (a[l] ^ v) * 0x100000001b3ull; a[l] ^= a[l] >> 29;
        b[l] = (b[l] + v) * 0xff51afd7ed558ccdull; b[l] ^= b[l] >> 32;
    }
    inline uint64_t fold(const uint64_t (&x)[4]) const {
        uint64_t r = x[0];
        for (int k = 1; k < 4; ++k) { r = (r ^ x[k]) * 0x9e3779b97f4a7c15ull; r ^= r >> 31; }
        return r;
    }
};
// returns whether the graph is decode-sized (every MUL_MAT has <= 8 activation columns): the same pass over the nodes answers both questions
bool graph_key_of(const ggml_cgraph * g, GraphKey & key) {
    KeyHasher H;
    const int nn = ggml_graph_n_nodes((ggml_cgraph *) g);
    bool small = true;
    for (int i = 0; i < nn; ++i) {
        const ggml_tensor * n = ggml_graph_node((ggml_cgraph *) g, i);
        if (n->op == GGML_OP_MUL_MAT && n->src[1]->ne[1] > 8) small = false;
        H.add((uint64_t) n->op | ((uint64_t) n->type << 32)); H.add((uint64_t) (uintptr_t) n->data); H.add((uint64_t) n->flags);
        for (int d = 0; d < 4; ++d) { H.add((uint64_t) n->ne[d]); H.add((uint64_t) n->nb[d]); }
        for (int s = 0; s < GGML_MAX_SRC; ++s) if (n->src[s]) {
            H.add((uint64_t) (uintptr_t) n->src[s]->data ^ ((uint64_t) s << 56));
            for (int d = 0; d < 4; ++d) { H.add((uint64_t) n->src[s]->ne[d]); H.add((uint64_t) n->src[s]->nb[d]); }
            H.add((uint64_t) n->src[s]->type);
        }
        const uint64_t * op64 = (const uint64_t *) n->op_params;
        for (size_t w = 0; w < GGML_MAX_OP_PARAMS / sizeof(uint64_t); ++w) H.add(op64[w]);
    }
    key.h1 = H.fold(H.a); key.h2 = H.fold(H.b); key.n = nn;
    return small;
}

enum ggml_status b200_backend_graph_compute_impl(ggml_backend_t backend, ggml_cgraph * g);
// GGML_B200_HOST_TIMING=1: host time spent inside graph_compute (issuing launches; the GPU runs asynchronously), reported per backend when it is freed
enum ggml_status b200_backend_graph_compute(ggml_backend_t backend, ggml_cgraph * g) {
    static const bool timing = getenv("GGML_B200_HOST_TIMING") && atoi(getenv("GGML_B200_HOST_TIMING")) != 0;
    if (!timing) return b200_backend_graph_compute_impl(backend, g);
    BackendCtx * c = (BackendCtx *) backend->context;
    const auto t0 = std::chrono::steady_clock::now();
    const enum ggml_status st = b200_backend_graph_compute_impl(backend, g);
    const long long dt = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    c->host_ns += dt; if (c->host_min_ns == 0 || dt < c->host_min_ns) c->host_min_ns = dt;
    c->host_calls += 1; c->host_nodes += ggml_graph_n_nodes(g);
    return st;
}
enum ggml_status b200_backend_graph_compute_impl(ggml_backend_t backend, ggml_cgraph * g) {
    BackendCtx * c = (BackendCtx *) backend->context;
    CUDA_OK(cudaSetDevice(c->device));
    const int nn = ggml_graph_n_nodes(g);
    // Decode-sized graphs (every MUL_MAT has <= 8 columns) are launch-bound: capture them once into a CUDA graph and replay while the
    // node list is unchanged (pointers, shapes and parameters; tensor CONTENTS may change) — like the reference's CUDA-graph path,
    // which is also limited to batch-1 graphs (ggml-cuda.cu:2725-2790).
    // host-overhead probe (profiles/): GGML_B200_NULL_COMPUTE=1 skips every launch of decode-sized graphs (2: of every graph), so `llama-bench -n` / `-p` then times
    // the reference's own host work (graph build, scheduling, the CPU-side token_embd GET_ROWS, input copies, logits read-back) plus this function's bookkeeping
    static const int null_compute = getenv("GGML_B200_NULL_COMPUTE") ? atoi(getenv("GGML_B200_NULL_COMPUTE")) : 0;
    GraphKey key;
    const bool small = c->graphs_enabled && nn >= 8 && graph_key_of(g, key);       // one pass: the replay key and "every MUL_MAT has <= 8 columns"
    if (!small) return null_compute >= 2 ? GGML_STATUS_SUCCESS : run_nodes(c, g);
    if (null_compute) return GGML_STATUS_SUCCESS;
    // the token after token case first: the same graph as last time on the whole-token engine needs nothing else from the host
    if (c->engine_enabled && c->engine && key == c->engine_key) {
        for (int idx : c->engine_pre) { int rc = 0; run_node(c, g, idx, rc); if (rc != B200_OK) return GGML_STATUS_FAILED; }
        const int rc = b200_decoder_step(c->engine, c->engine_n_kv, c->stream);
        if (rc != B200_OK) { B200_LOG("decoder step failed: %s", b200_error_string(rc)); return GGML_STATUS_FAILED; }
        return GGML_STATUS_SUCCESS;
    }
    size_t need = 0;
    for (int i = 0; i < nn; ++i) {
        const ggml_tensor * n = ggml_graph_node(g, i);
        const size_t sb = is_noop(n) ? 0 : node_scratch_bytes(n);
        if (sb > need) need = sb;
    }
    scratch_for(c, need);                                            // no allocation may happen while capturing
    auto run_pre = [&](const std::vector<int> & pre) {                // the engine's per-op prefix, re-resolved from THIS graph by node index
        for (int idx : pre) { int rc = 0; run_node(c, g, idx, rc); if (rc != B200_OK) return false; }
        return true;
    };
    // ---- whole-token decode engine: one persistent kernel for the split when it is the llama-family batch-1 decoder pattern
    if (c->engine_enabled) {
        if (c->engine && key == c->engine_key) {
            if (!run_pre(c->engine_pre)) return GGML_STATUS_FAILED;
            const int rc = b200_decoder_step(c->engine, c->engine_n_kv, c->stream);
            if (rc != B200_OK) { B200_LOG("decoder step failed: %s", b200_error_string(rc)); return GGML_STATUS_FAILED; }
            return GGML_STATUS_SUCCESS;
        }
        if (key != c->engine_reject_key) {
            DecoderMatch M;
            void * h = nullptr;
            if (match_decoder(g, M) && b200_decoder_create(&M.d, &h) == B200_OK) {
                if (c->engine) { CUDA_OK(cudaStreamSynchronize(c->stream)); b200_decoder_destroy(c->engine); }
                c->engine = h; c->engine_n_kv = M.n_kv; c->engine_pre = M.pre; c->engine_key = key;
                if (!run_pre(c->engine_pre)) return GGML_STATUS_FAILED;
                const int rc = b200_decoder_step(c->engine, c->engine_n_kv, c->stream);
                if (rc != B200_OK) { B200_LOG("decoder step failed: %s", b200_error_string(rc)); return GGML_STATUS_FAILED; }
                return GGML_STATUS_SUCCESS;
            }
            c->engine_reject_key = key;                              // not the pattern (or unsupported types): remember, use the per-op path
        }
    }
    if (c->graph_exec && key == c->graph_key) {
        CUDA_OK(cudaGraphLaunch(c->graph_exec, c->stream));
        return GGML_STATUS_SUCCESS;
    }
    // a node list seen for the first time runs eagerly (module loading, one-time attribute setup); it is captured when it comes back
    if (key != c->pending_key) { c->pending_key = key; return run_nodes(c, g); }
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; c->graph_key.clear(); }
    cudaGraph_t graph = nullptr;
    CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const enum ggml_status status = run_nodes(c, g);
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (status != GGML_STATUS_SUCCESS || e != cudaSuccess || !graph) {
        (void) cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        c->graphs_enabled = false;                                    // something on this path is not capturable: plain launches from now on
        return status != GGML_STATUS_SUCCESS ? status : run_nodes(c, g);
    }
    e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { (void) cudaGetLastError(); c->graph_exec = nullptr; c->graphs_enabled = false; return run_nodes(c, g); }
    c->graph_key = key;
    CUDA_OK(cudaGraphLaunch(c->graph_exec, c->stream));
    return GGML_STATUS_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------------- backend (stream)
const char * b200_backend_get_name(ggml_backend_t backend) { return ((BackendCtx *) backend->context)->name.c_str(); }

void b200_backend_free(ggml_backend_t backend) {
    BackendCtx * c = (BackendCtx *) backend->context;
    if (c->host_calls) B200_LOG("%s: synchronize() called %lld times, host blocked %.3f ms in them", c->name.c_str(), c->sync_calls, c->sync_ns / 1e6);
    if (c->host_calls) B200_LOG("%s: host time inside graph_compute: %.3f ms over %lld calls (%lld nodes): %.1f us per call (fastest call %.1f us), %.2f us per node", c->name.c_str(), c->host_ns / 1e6,
                                c->host_calls, c->host_nodes, c->host_ns / 1e3 / c->host_calls, c->host_min_ns / 1e3, c->host_ns / 1e3 / (c->host_nodes ? c->host_nodes : 1));
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->engine) b200_decoder_destroy(c->engine);
    if (c->scratch) cudaFree(c->scratch);
    if (c->hoist_buf) cudaFree(c->hoist_buf);
    if (c->hop_event) cudaEventDestroy(c->hop_event);
    cudaStreamDestroy(c->stream);
    delete c;
    delete backend;
}

void b200_backend_set_tensor_async(ggml_backend_t backend, ggml_tensor * tensor, const void * data, size_t offset, size_t size) {
    BackendCtx * c = (BackendCtx *) backend->context;
    if (is_planar(tensor)) { cudaStreamSynchronize(c->stream); b200_buf_set_tensor(tensor->buffer, tensor, data, offset, size); return; }
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaMemcpyAsync((char *) tensor->data + offset, data, size, cudaMemcpyHostToDevice, c->stream));
}
void b200_backend_get_tensor_async(ggml_backend_t backend, const ggml_tensor * tensor, void * data, size_t offset, size_t size) {
    BackendCtx * c = (BackendCtx *) backend->context;
    if (is_planar(tensor)) { cudaStreamSynchronize(c->stream); b200_buf_get_tensor(tensor->buffer, tensor, data, offset, size); return; }
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaMemcpyAsync(data, (const char *) tensor->data + offset, size, cudaMemcpyDeviceToHost, c->stream));
}

bool backend_is_b200(ggml_backend_t b) { return b && b->iface.get_name == b200_backend_get_name; }

// layer-split pipeline hop: the hidden state crosses a device boundary as ONE peer copy ordered on both streams.
// (bench.py --gpus N uses NCCL send/recv for the same hop between its one-process-per-GPU ranks; inside ONE process — which is how the
// reference's scheduler drives several devices, ggml-backend.cpp:1539 — a peer copy over NVLink is the native primitive.)
bool b200_backend_cpy_tensor_async(ggml_backend_t src_backend, ggml_backend_t dst_backend, const ggml_tensor * src, ggml_tensor * dst) {
    if (!backend_is_b200(src_backend) || !backend_is_b200(dst_backend)) return false;
    if (!src->buffer || !dst->buffer || !buffer_is_b200(src->buffer) || !buffer_is_b200(dst->buffer)) return false;
    if (is_planar(src) || is_planar(dst) || ggml_nbytes(src) != ggml_nbytes(dst) || !ggml_is_contiguous(src) || !ggml_is_contiguous(dst)) return false;
    BackendCtx * sc = (BackendCtx *) src_backend->context, * dc = (BackendCtx *) dst_backend->context;
    if (sc == dc) {
        CUDA_OK(cudaSetDevice(sc->device));
        CUDA_OK(cudaMemcpyAsync(dst->data, src->data, ggml_nbytes(src), cudaMemcpyDeviceToDevice, sc->stream));
        return true;
    }
    // copy on the SOURCE stream (after the producer), then make the destination stream wait for it
    CUDA_OK(cudaSetDevice(sc->device));
    if (sc->device == dc->device) CUDA_OK(cudaMemcpyAsync(dst->data, src->data, ggml_nbytes(src), cudaMemcpyDeviceToDevice, sc->stream));
    else CUDA_OK(cudaMemcpyPeerAsync(dst->data, dc->device, src->data, sc->device, ggml_nbytes(src), sc->stream));
    if (!sc->hop_event) CUDA_OK(cudaEventCreateWithFlags(&sc->hop_event, cudaEventDisableTiming));      // one event per source backend, re-recorded per hop
    CUDA_OK(cudaEventRecord(sc->hop_event, sc->stream));
    CUDA_OK(cudaSetDevice(dc->device));
    CUDA_OK(cudaStreamWaitEvent(dc->stream, sc->hop_event, 0));      // the wait captures the record above; a later re-record does not move it
    return true;
}

void b200_backend_synchronize(ggml_backend_t backend) {
    BackendCtx * c = (BackendCtx *) backend->context;
    CUDA_OK(cudaSetDevice(c->device));
    static const bool timing = getenv("GGML_B200_HOST_TIMING") && atoi(getenv("GGML_B200_HOST_TIMING")) != 0;
    const auto t0 = timing ? std::chrono::steady_clock::now() : std::chrono::steady_clock::time_point();
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (timing) { c->sync_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); c->sync_calls += 1; }
}

void b200_backend_event_record(ggml_backend_t backend, ggml_backend_event_t event) {
    BackendCtx * c = (BackendCtx *) backend->context;
    CUDA_OK(cudaEventRecord((cudaEvent_t) event->context, c->stream));
}
void b200_backend_event_wait(ggml_backend_t backend, ggml_backend_event_t event) {
    BackendCtx * c = (BackendCtx *) backend->context;
    CUDA_OK(cudaStreamWaitEvent(c->stream, (cudaEvent_t) event->context, 0));
}

const ggml_backend_i b200_backend_iface = {
    /* get_name           */ b200_backend_get_name,
    /* free               */ b200_backend_free,
    /* set_tensor_async   */ b200_backend_set_tensor_async,
    /* get_tensor_async   */ b200_backend_get_tensor_async,
    /* cpy_tensor_async   */ b200_backend_cpy_tensor_async,
    /* synchronize        */ b200_backend_synchronize,
    /* graph_plan_create  */ nullptr,
    /* graph_plan_free    */ nullptr,
    /* graph_plan_update  */ nullptr,
    /* graph_plan_compute */ nullptr,
    /* graph_compute      */ b200_backend_graph_compute,
    /* event_record       */ b200_backend_event_record,
    /* event_wait         */ b200_backend_event_wait,
    /* graph_optimize     */ nullptr,
};

ggml_guid_t b200_guid() {
    static ggml_guid guid = { 0xb2, 0x00, 0x5a, 0x10, 0x0a, 0x67, 0x67, 0x6d, 0x6c, 0x2d, 0x62, 0x32, 0x30, 0x30, 0x00, 0x01 };
    return &guid;
}

// ---------------------------------------------------------------------------------------------------------------- device
const char * b200_dev_get_name(ggml_backend_dev_t dev) { return ((DeviceCtx *) dev->context)->name.c_str(); }
const char * b200_dev_get_description(ggml_backend_dev_t dev) { return ((DeviceCtx *) dev->context)->description.c_str(); }
void b200_dev_get_memory(ggml_backend_dev_t dev, size_t * free, size_t * total) {
    CUDA_OK(cudaSetDevice(((DeviceCtx *) dev->context)->index));
    CUDA_OK(cudaMemGetInfo(free, total));
}
enum ggml_backend_dev_type b200_dev_get_type(ggml_backend_dev_t) { return GGML_BACKEND_DEVICE_TYPE_GPU; }
void b200_dev_get_props(ggml_backend_dev_t dev, ggml_backend_dev_props * props) {
    DeviceCtx * d = (DeviceCtx *) dev->context;
    memset(props, 0, sizeof(*props));
    props->name = d->name.c_str(); props->description = d->description.c_str(); props->type = GGML_BACKEND_DEVICE_TYPE_GPU;
    props->device_id = d->pci_id.empty() ? nullptr : d->pci_id.c_str();      // used by llama.cpp to de-duplicate devices (llama.cpp:208-229)
    b200_dev_get_memory(dev, &props->memory_free, &props->memory_total);
    props->caps.async = true; props->caps.host_buffer = true; props->caps.buffer_from_host_ptr = false; props->caps.events = true;
}

ggml_backend_t b200_dev_init_backend(ggml_backend_dev_t dev, const char *) {
    DeviceCtx * d = (DeviceCtx *) dev->context;
    if (cudaSetDevice(d->index) != cudaSuccess) return nullptr;
    BackendCtx * c = new BackendCtx(); c->device = d->index; c->name = d->name;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return nullptr; }
    if (const char * e = getenv("GGML_B200_DISABLE_GRAPHS")) c->graphs_enabled = atoi(e) == 0;
    if (const char * e = getenv("GGML_B200_DISABLE_ENGINE")) c->engine_enabled = atoi(e) == 0;
    ggml_backend * b = new ggml_backend;
    b->guid = b200_guid(); b->iface = b200_backend_iface; b->device = dev; b->context = c;
    return b;
}
ggml_backend_buffer_type_t b200_dev_get_buffer_type(ggml_backend_dev_t dev) { return &((DeviceCtx *) dev->context)->buft; }
bool b200_dev_supports_buft(ggml_backend_dev_t dev, ggml_backend_buffer_type_t buft) {
    // (the pinned host type is for the CPU backend's tensors: our kernels take device pointers only, so it is NOT a buffer type we compute from)
    return buft->iface.get_name == b200_buft_get_name && buft->device == dev;
}
bool b200_dev_offload_op(ggml_backend_dev_t, const ggml_tensor * op) {
    // weights left on the host: worth shipping to the GPU only for real batches (same threshold as ggml-cuda.cu:3724-3731)
    const int min_batch = 32;
    return (op->ne[1] >= min_batch && op->op != GGML_OP_GET_ROWS) || (op->ne[2] >= min_batch && op->op == GGML_OP_MUL_MAT_ID);
}
ggml_backend_event_t b200_dev_event_new(ggml_backend_dev_t dev) {
    CUDA_OK(cudaSetDevice(((DeviceCtx *) dev->context)->index));
    cudaEvent_t ev;
    CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    return new ggml_backend_event{ dev, ev };
}
void b200_dev_event_free(ggml_backend_dev_t, ggml_backend_event_t event) { cudaEventDestroy((cudaEvent_t) event->context); delete event; }
void b200_dev_event_synchronize(ggml_backend_dev_t, ggml_backend_event_t event) { CUDA_OK(cudaEventSynchronize((cudaEvent_t) event->context)); }

const ggml_backend_device_i b200_device_iface = {
    /* get_name             */ b200_dev_get_name,
    /* get_description      */ b200_dev_get_description,
    /* get_memory           */ b200_dev_get_memory,
    /* get_type             */ b200_dev_get_type,
    /* get_props            */ b200_dev_get_props,
    /* init_backend         */ b200_dev_init_backend,
    /* get_buffer_type      */ b200_dev_get_buffer_type,
    /* get_host_buffer_type */ b200_dev_get_host_buffer_type,
    /* buffer_from_host_ptr */ nullptr,
    /* supports_op          */ b200_dev_supports_op,
    /* supports_buft        */ b200_dev_supports_buft,
    /* offload_op           */ b200_dev_offload_op,
    /* event_new            */ b200_dev_event_new,
    /* event_free           */ b200_dev_event_free,
    /* event_synchronize    */ b200_dev_event_synchronize,
};

// ---------------------------------------------------------------------------------------------------------------- registry
int probe_devices_once() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void) cudaGetLastError(); n = 0; }
    int kept = 0;
    for (int i = 0; i < n && kept < MAX_DEVICES; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) != cudaSuccess) continue;
        if (p.major != 10) { B200_LOG("device %d (%s, sm_%d%d) skipped: kernels are built for sm_100a only", i, p.name, p.major, p.minor); continue; }
        DeviceCtx & d = g_devices[kept];
        d.index = i; d.name = "B200:" + std::to_string(kept); d.description = p.name;
        char pci[32]; snprintf(pci, sizeof(pci), "%04x:%02x:%02x.0", p.pciDomainID, p.pciBusID, p.pciDeviceID); d.pci_id = pci;
        d.buft.iface = { b200_buft_get_name, b200_buft_alloc, b200_buft_alignment, b200_buft_max_size, b200_buft_alloc_size, b200_buft_is_host };
        d.buft.device = &d.dev; d.buft.context = &d;
        d.dev.iface = b200_device_iface; d.dev.reg = &g_reg; d.dev.context = &d;
        ++kept;
    }
    // peers: the layer-split hop is a direct NVLink copy
    for (int a = 0; a < kept; ++a) for (int b = 0; b < kept; ++b) if (a != b) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, g_devices[a].index, g_devices[b].index) == cudaSuccess && can) {
            cudaSetDevice(g_devices[a].index);
            cudaError_t e = cudaDeviceEnablePeerAccess(g_devices[b].index, 0);
            if (e != cudaSuccess) (void) cudaGetLastError();
        }
    }
    g_n_devices = kept;
    return kept;
}
int probe_devices() {                                                     // any thread, any entry point (omni creates its backends from 3 threads)
    static std::once_flag once;
    std::call_once(once, [] { probe_devices_once(); });
    return g_n_devices;
}

const char * b200_reg_get_name(ggml_backend_reg_t) { return "B200"; }
size_t b200_reg_get_device_count(ggml_backend_reg_t) { return (size_t) probe_devices(); }
ggml_backend_dev_t b200_reg_get_device(ggml_backend_reg_t, size_t index) {
    return index < (size_t) probe_devices() ? &g_devices[index].dev : nullptr;
}
void * b200_reg_get_proc_address(ggml_backend_reg_t, const char * name) {
    (void) name;        // optional hooks consumers ask for (split buffer type, set_n_threads, features, host-buffer registration): none
    return nullptr;
}

} // namespace

extern "C" {

ggml_backend_reg_t ggml_backend_b200_reg(void) {
    static std::once_flag once;
    std::call_once(once, [] {
        g_reg.api_version = GGML_BACKEND_API_VERSION;
        g_reg.iface = { b200_reg_get_name, b200_reg_get_device_count, b200_reg_get_device, b200_reg_get_proc_address };
        g_reg.context = nullptr;
        probe_devices();
    });
    return &g_reg;
}

ggml_backend_reg_t ggml_backend_init(void) { return ggml_backend_b200_reg(); }

// 0 = cannot run here (no sm_100 device): the loader then skips the library (ggml-backend-reg.cpp:257-273)
int ggml_backend_score(void) { ggml_backend_b200_reg(); return g_n_devices > 0 ? 100 : 0; }

ggml_backend_t ggml_backend_b200_init(int device) {
    ggml_backend_b200_reg();
    if (device < 0 || device >= g_n_devices) return nullptr;
    return b200_dev_init_backend(&g_devices[device].dev, nullptr);
}

int ggml_backend_b200_abi(void) { return b200_abi_version(); }

} // extern "C"
