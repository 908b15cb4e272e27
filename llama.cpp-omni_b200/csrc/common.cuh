// common.cuh — shared device/host helpers for the sm_100a kernel library (libb200ops.so).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/b200_ops.h"

#define B200_CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return -(int) e_; } while (0)
#define B200_LAUNCH_CHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return -(int) e_; } while (0)

namespace b200 {

constexpr int WARP = 32;

// ---- block geometry of the ggml quantised types (ggml-common.h:170-336) -------------------------------------------
template <int T> struct qtraits;
template <> struct qtraits<B200_Q4_0> { static constexpr int qk = 32,  bytes = 18,  payload = 16,  d_off = 0,   p_off = 2;  };
template <> struct qtraits<B200_Q8_0> { static constexpr int qk = 32,  bytes = 34,  payload = 32,  d_off = 0,   p_off = 2;  };
template <> struct qtraits<B200_Q4_K> { static constexpr int qk = 256, bytes = 144, payload = 144, d_off = 0,   p_off = 0;  };
template <> struct qtraits<B200_Q5_K> { static constexpr int qk = 256, bytes = 176, payload = 176, d_off = 0,   p_off = 0;  };
template <> struct qtraits<B200_Q6_K> { static constexpr int qk = 256, bytes = 210, payload = 208, d_off = 208, p_off = 0;  };

__host__ __device__ inline bool is_quant(int t) { return t == B200_Q4_0 || t == B200_Q8_0 || t == B200_Q4_K || t == B200_Q5_K || t == B200_Q6_K; }
__host__ __device__ inline bool is_kquant(int t) { return t == B200_Q4_K || t == B200_Q5_K || t == B200_Q6_K; }
__host__ __device__ inline int  blck_size(int t) { return is_kquant(t) ? 256 : (t == B200_Q4_0 || t == B200_Q8_0) ? 32 : 1; }
__host__ __device__ inline int  type_size(int t) {
    switch (t) { case B200_F32: case B200_I32: return 4; case B200_F16: case B200_BF16: return 2; case B200_I64: return 8;
                 case B200_Q4_0: return 18; case B200_Q8_0: return 34; case B200_Q4_K: return 144; case B200_Q5_K: return 176;
                 case B200_Q6_K: return 210; }
    return 0;
}
__host__ __device__ inline int  payload_size(int t) { return t == B200_Q4_0 ? 16 : t == B200_Q8_0 ? 32 : t == B200_Q6_K ? 208 : type_size(t); }

// ---- activation record (see b200_ops.h) ---------------------------------------------------------------------------
// G = 256 : qs[k] | float d[k/256] | int16 bsum[k/16]        (q8_K semantics)
// G = 32  : qs[k] | half  d[k/32]  | int16 bsum[k/32]        (q8_0 semantics)
struct ActLayout {
    int64_t k; int group; int64_t d_off, bsum_off, bytes;
};
__host__ __device__ inline ActLayout act_layout(int weight_type, int64_t k) {
    ActLayout L; L.k = k;
    if (is_kquant(weight_type)) { L.group = 256; L.d_off = k; L.bsum_off = k + 4*(k/256); L.bytes = L.bsum_off + 2*(k/16); }
    else                        { L.group = 32;  L.d_off = k; L.bsum_off = k + 2*(k/32);  L.bytes = L.bsum_off + 2*(k/32); }
    L.bytes = (L.bytes + 15) & ~(int64_t) 15;
    return L;
}

// ---- loads ----------------------------------------------------------------------------------------------------------
// streaming 128-bit load for weights: read-only path, do not allocate in L1 (weights are touched exactly once per token)
__device__ __forceinline__ uint4 ldg_stream16(const void * p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// alignment-agnostic 16-byte load (native q4_0/q8_0/q6_K blocks are only 2-byte aligned)
__device__ __forceinline__ uint4 ld16_any(const uint8_t * p) {
    const uintptr_t a = (uintptr_t) p;
    uint4 r;
    if ((a & 15) == 0) { r = __ldg((const uint4 *) p); }
    else if ((a & 3) == 0) { const uint32_t * q = (const uint32_t *) p; r.x = __ldg(q); r.y = __ldg(q + 1); r.z = __ldg(q + 2); r.w = __ldg(q + 3); }
    else { const uint16_t * q = (const uint16_t *) p;
           r.x = __ldg(q)     | ((uint32_t) __ldg(q + 1) << 16); r.y = __ldg(q + 2) | ((uint32_t) __ldg(q + 3) << 16);
           r.z = __ldg(q + 4) | ((uint32_t) __ldg(q + 5) << 16); r.w = __ldg(q + 6) | ((uint32_t) __ldg(q + 7) << 16); }
    return r;
}
template <bool ALIGNED> __device__ __forceinline__ uint4 ld16_w(const uint8_t * p) {
    if constexpr (ALIGNED) return ldg_stream16(p); else return ld16_any(p);
}
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// cudaFuncSetAttribute is PER DEVICE: a process that drives several GPUs (the ggml scheduler's layer split) must raise the dynamic
// shared-memory limit of a kernel on each of them.  `mask` is a per-kernel static bit set of devices already done.
// Several host threads may drive backends of one process at once (omni's LLM / TTS / encoder threads): the mask is atomic, and two threads racing on
// the same device at worst both call cudaFuncSetAttribute, which is idempotent.
typedef std::atomic<unsigned long long> smem_mask_t;
template <class K> inline cudaError_t ensure_dyn_smem(K kernel, int bytes, smem_mask_t & mask) {
    int dev = 0; cudaGetDevice(&dev);
    if (mask.load(std::memory_order_acquire) >> (dev & 63) & 1ull) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
    return e;
}

inline int sm_count() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}


// ---- prepared activations of the tensor-core GEMMs (mmq_tc.cu): F16, tiled [row / 256][k / 64][32 KB], each tile in the canonical K-major UMMA layout (8-row x
// 16-byte core matrices; K direction 4 KB apart, row groups 128 B apart).  Producers that feed a MUL_MAT (RMS_NORM, GLU, FLASH_ATTN_EXT) can write this layout
// directly instead of F32 + a conversion pass.  Byte offset of the 16-byte core-matrix row that holds columns [col, col + 8) of `row` (col % 8 == 0):
__host__ __device__ inline int64_t act_tile_off(int64_t row, int64_t col, int64_t k) {
    return ((row >> 8) * (k >> 6) + (col >> 6)) * 32768 + ((col & 63) >> 3) * 4096 + ((row & 255) >> 3) * 128 + (row & 7) * 16;
}

} // namespace b200
