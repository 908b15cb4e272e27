// repack.cu — native <-> planar weight layout (see include/b200_ops.h "Weight layouts in HBM").
//
// The reference CUDA backend keeps ggml's block structs verbatim and reads q4_0 / q8_0 / q6_K with 2- and 4-byte loads
// (vecdotq.cuh:103-122, 580-600) because those blocks are 18 / 34 / 210 bytes.  The buffer interface makes the device
// layout private (ggml-backend-impl.h:41-58), so — like the CPU backend's CPU_REPACK buffer type (ggml-cpu/repack.cpp:1872) —
// this backend re-lays such tensors at upload: a 16-byte aligned payload plane followed by an f16 `d` plane.  Uploads may
// arrive in arbitrary byte chunks (llama-model-loader.cpp:1077-1093), so both directions work on (offset, size) ranges.
#include "common.cuh"

namespace b200 {

template <bool SCATTER>
__global__ void __launch_bounds__(256) k_repack(const uint8_t * __restrict__ src, uint8_t * __restrict__ dst, int64_t nblocks,
                                                int64_t offset, int64_t size, int bytes, int payload, int d_off, int p_off) {
    // SCATTER: src = native chunk (src[i] is native byte offset+i), dst = planar base.  GATHER: src = planar base, dst = native chunk.
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (int64_t) gridDim.x * blockDim.x) {
        const int64_t nat = offset + i, b = nat / bytes; const int o = (int) (nat % bytes);
        const int64_t pl = (o >= d_off && o < d_off + 2) ? nblocks * payload + b * 2 + (o - d_off) : b * payload + (o - p_off);
        if (SCATTER) dst[pl] = src[i]; else dst[i] = src[pl];
    }
}

static int repack_run(bool scatter, int type, const void * src, void * dst, int64_t nblocks, int64_t offset, int64_t size, cudaStream_t st) {
    if (!(type == B200_Q4_0 || type == B200_Q8_0 || type == B200_Q6_K)) return B200_ERR_UNSUPPORTED;
    const int bytes = type_size(type), payload = payload_size(type);
    const int d_off = type == B200_Q6_K ? 208 : 0, p_off = type == B200_Q6_K ? 0 : 2;
    if (offset < 0 || size < 0 || offset + size > nblocks * bytes) return B200_ERR_ARG;
    if (size == 0) return B200_OK;
    int64_t g = (size + 255) / 256; const int64_t cap = (int64_t) sm_count() * 32; if (g > cap) g = cap;
    if (scatter) k_repack<true ><<<(unsigned) g, 256, 0, st>>>((const uint8_t *) src, (uint8_t *) dst, nblocks, offset, size, bytes, payload, d_off, p_off);
    else         k_repack<false><<<(unsigned) g, 256, 0, st>>>((const uint8_t *) src, (uint8_t *) dst, nblocks, offset, size, bytes, payload, d_off, p_off);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

} // namespace b200

using namespace b200;

extern "C" int b200_repack_supported(int type) { return type == B200_Q4_0 || type == B200_Q8_0 || type == B200_Q6_K; }
extern "C" int b200_repack_scatter(int type, const void * src_native_chunk, void * dst_planar, int64_t nblocks_total, int64_t offset,
                                   int64_t size, void * stream) {
    return repack_run(true, type, src_native_chunk, dst_planar, nblocks_total, offset, size, (cudaStream_t) stream);
}
extern "C" int b200_repack_gather(int type, const void * src_planar, void * dst_native_chunk, int64_t nblocks_total, int64_t offset,
                                  int64_t size, void * stream) {
    return repack_run(false, type, src_planar, dst_native_chunk, nblocks_total, offset, size, (cudaStream_t) stream);
}

// ---- library / device ---------------------------------------------------------------------------------------------------
extern "C" int b200_abi_version(void) { return 3; }          // 3: round 2 (mul_mat_add / _multi / _glu, unary_param, the Token2Wav ops, ...)
extern "C" const char * b200_error_string(int code) {
    if (code == B200_OK) return "ok";
    if (code == B200_ERR_UNSUPPORTED) return "b200: unsupported type/shape for this kernel";
    if (code == B200_ERR_ARG) return "b200: bad argument";
    if (code < 0) return cudaGetErrorString((cudaError_t) (-code));
    return "b200: unknown";
}
extern "C" int b200_device_sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return n;
}
