// hop.cu — the layer-split pipeline hop ON THE DEVICE: the hidden state of a token crosses a stage boundary as ONE peer write
// over NVLink into the next stage's input buffer, followed by a release flag; the next stage's stream waits for the flag with a 1-CTA kernel in
// front of its decode step.  No host code runs between two stages' kernels, so a stage's step (wait -> mask cast -> k_stream -> send) is one CUDA
// graph replay per token.  Replaces, for one-process-per-GPU launches, the send/recv pair the reference issues per split boundary
// (ggml_backend_sched_compute_splits -> cpy_tensor_async, ggml/src/ggml-backend.cpp:1539; ggml-cuda.cu:2598-2620 cudaMemcpyPeerAsync + event).
//
// Link protocol (producer P -> consumer C, `n_slots` buffers used round-robin so P can run one step ahead):
//   C owns   data[slot][n] (f32)  and  ready[slot] (u32 sequence number, written by P)
//   P owns   ack[slot]            (u32 sequence number, written by C once its step that READ the slot has finished)
//   send:  seq = ++count; wait ack[slot] >= seq - 1; copy; __threadfence_system; ready[slot] = seq   (st.release.sys)
//   wait:  seq = ++count; wait ready[slot] >= seq                                                     (ld.acquire.sys)
//   ack :  ack[slot] = seq of the step that just finished                                              (st.release.sys)
// Sequence counters live in device memory, so a captured graph can be replayed without host help.  Every wait is bounded (~2 s): on time-out
// the kernel raises the link's error word instead of hanging the device.
#include "common.cuh"
#include <string.h>

namespace b200 {

struct HopState { unsigned count; unsigned error; };

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned * p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_sys(unsigned * p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

__device__ __forceinline__ bool hop_spin(const unsigned * flag, unsigned want) {
    const long long t0 = clock64();
    while ((int) (ld_acquire_sys(flag) - want) < 0) {
        if (clock64() - t0 > 4000000000ll) return false;
        __nanosleep(40);
    }
    return true;
}

// one CTA: wait for the consumer's ack of this slot's previous tenant, copy n floats into the peer buffer, publish
__global__ void k_hop_send(const float * __restrict__ src, float * __restrict__ dst_peer, int n, unsigned * ready_peer, const unsigned * ack_local, HopState * st) {
    __shared__ unsigned s_seq; __shared__ int s_ok;
    if (threadIdx.x == 0) {
        const unsigned seq = st->count + 1; st->count = seq; s_seq = seq;
        s_ok = hop_spin(ack_local, seq - 1) ? 1 : 0;
        if (!s_ok) st->error = 1;
    }
    __syncthreads();
    if (!s_ok) return;
    for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4) *(float4 *) (dst_peer + i) = *(const float4 *) (src + i);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(ready_peer, s_seq);
}

// one CTA in front of the stage's decode step: the step's kernels are stream-ordered behind it
__global__ void k_hop_wait(const unsigned * ready_local, HopState * st) {
    if (threadIdx.x == 0) {
        const unsigned seq = st->count + 1; st->count = seq;
        if (!hop_spin(ready_local, seq)) st->error = 2;
    }
}

// after the stage's step: tell the producer that the slot it wrote has been consumed
__global__ void k_hop_ack(unsigned * ack_peer, HopState * st) {
    if (threadIdx.x == 0) { const unsigned seq = st->count + 1; st->count = seq; st_release_sys(ack_peer, seq); }
}

} // namespace b200

using namespace b200;

// ---- memory that another process on the node can map (cudaIpc*): the consumer's input slots + flags, the producer's ack words -----------------
extern "C" int b200_ipc_alloc(size_t bytes, void ** ptr, void * handle64) {
    if (!ptr || !handle64 || bytes == 0) return B200_ERR_ARG;
    B200_CUDA_TRY(cudaMalloc(ptr, bytes));
    B200_CUDA_TRY(cudaMemset(*ptr, 0, bytes));
    B200_CUDA_TRY(cudaDeviceSynchronize());
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    B200_CUDA_TRY(cudaIpcGetMemHandle((cudaIpcMemHandle_t *) handle64, *ptr));
    return B200_OK;
}
extern "C" int b200_ipc_open(const void * handle64, void ** ptr) {
    if (!ptr || !handle64) return B200_ERR_ARG;
    cudaIpcMemHandle_t h; memcpy(&h, handle64, sizeof(h));
    B200_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return B200_OK;
}
extern "C" int b200_ipc_close(void * ptr) { if (ptr) B200_CUDA_TRY(cudaIpcCloseMemHandle(ptr)); return B200_OK; }
extern "C" int b200_ipc_free(void * ptr) { if (ptr) B200_CUDA_TRY(cudaFree(ptr)); return B200_OK; }

// state = 8 bytes of zero-initialised device memory per call site (sequence counter + error word)
extern "C" int b200_hop_send(const float * src, float * dst_peer, int64_t n, unsigned * ready_peer, const unsigned * ack_local, void * state, void * stream) {
    if (!src || !dst_peer || !ready_peer || !ack_local || !state || n <= 0 || n % 4 || ((uintptr_t) src | (uintptr_t) dst_peer) % 16) return B200_ERR_ARG;
    k_hop_send<<<1, 256, 0, (cudaStream_t) stream>>>(src, dst_peer, (int) n, ready_peer, ack_local, (HopState *) state);
    B200_CUDA_TRY(cudaGetLastError());
    return B200_OK;
}
extern "C" int b200_hop_wait(const unsigned * ready_local, void * state, void * stream) {
    if (!ready_local || !state) return B200_ERR_ARG;
    k_hop_wait<<<1, 32, 0, (cudaStream_t) stream>>>(ready_local, (HopState *) state);
    B200_CUDA_TRY(cudaGetLastError());
    return B200_OK;
}
extern "C" int b200_hop_ack(unsigned * ack_peer, void * state, void * stream) {
    if (!ack_peer || !state) return B200_ERR_ARG;
    k_hop_ack<<<1, 32, 0, (cudaStream_t) stream>>>(ack_peer, (HopState *) state);
    B200_CUDA_TRY(cudaGetLastError());
    return B200_OK;
}
