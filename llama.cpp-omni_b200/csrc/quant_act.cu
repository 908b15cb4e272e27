// quant_act.cu — activation quantisation for the decode matvec (replaces quantize_q8_1, ggml-cuda/quantize.cu:4-48).
//
// Arithmetic follows the CPU oracle bit for bit so that the integer dot products of the matvec match the reference:
//   K-quant weights  -> q8_K : quantize_row_q8_K_ref, ggml-quants.c:2555-2592  (iscale = -127/max, nearest-even, clamp 127)
//   q4_0/q8_0 weights-> q8_0 : quantize_row_q8_0 (x86 build), ggml-cpu/arch/x86/quants.c:297-360 (id = 127/amax, nearest-even)
// Output: the planar activation record described in include/b200_ops.h.
#include "quant_dev.cuh"

namespace b200 {

// one warp per 256-element super-block
__global__ void __launch_bounds__(256) k_quantize_q8K(const float * __restrict__ x, int64_t x_col_stride, uint8_t * __restrict__ act,
                                                      int64_t k, int64_t act_bytes, int64_t d_off, int64_t bsum_off) {
    const int lane = threadIdx.x & 31;
    const int64_t blk = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t col = blockIdx.y;
    if (blk * 256 >= k) return;
    const float * xs = x + col * x_col_stride + blk * 256 + lane * 8;
    const float4 v0 = *(const float4 *) xs, v1 = *(const float4 *) (xs + 4);
    const float v[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
    quant_block_q8K(v, act + col * act_bytes, blk, d_off, bsum_off);
}

// 8 lanes per 32-element block, 4 blocks per warp
__global__ void __launch_bounds__(256) k_quantize_q8_0(const float * __restrict__ x, int64_t x_col_stride, uint8_t * __restrict__ act,
                                                       int64_t k, int64_t act_bytes, int64_t d_off, int64_t bsum_off) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t col = blockIdx.y;
    const int64_t blk = warp * 4 + (lane >> 3);
    const bool live = blk * 32 < k;
    float4 v = make_float4(0, 0, 0, 0);
    if (live) v = *(const float4 *) (x + col * x_col_stride + blk * 32 + (lane & 7) * 4);
    quant_block_q8_0(v, live, act + col * act_bytes, blk, d_off, bsum_off);
}

} // namespace b200

using namespace b200;

extern "C" size_t b200_act_bytes(int weight_type, int64_t k) { return (size_t) act_layout(weight_type, k).bytes; }

extern "C" int b200_quantize_act(int weight_type, const float * x, int64_t x_col_stride, void * act, int64_t k, int64_t ncols,
                                 void * stream) {
    if (!is_quant(weight_type) || k <= 0 || k % blck_size(weight_type) != 0 || ncols <= 0) return B200_ERR_UNSUPPORTED;
    if (((uintptr_t) x & 15) || (x_col_stride & 3)) return B200_ERR_UNSUPPORTED;
    const ActLayout L = act_layout(weight_type, k);
    cudaStream_t st = (cudaStream_t) stream;
    if (L.group == 256) {
        const int64_t nblk = k / 256;
        dim3 grid((unsigned) ((nblk + 7) / 8), (unsigned) ncols);
        k_quantize_q8K<<<grid, 256, 0, st>>>(x, x_col_stride, (uint8_t *) act, k, L.bytes, L.d_off, L.bsum_off);
    } else {
        const int64_t nwarp = (k / 32 + 3) / 4;
        dim3 grid((unsigned) ((nwarp + 7) / 8), (unsigned) ncols);
        k_quantize_q8_0<<<grid, 256, 0, st>>>(x, x_col_stride, (uint8_t *) act, k, L.bytes, L.d_off, L.bsum_off);
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}
