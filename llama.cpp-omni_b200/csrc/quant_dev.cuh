// quant_dev.cuh — device-side activation quantisers shared by quant_act.cu (stand-alone launch) and the fused producers.
// Arithmetic = the CPU oracle's, bit for bit (see quant_act.cu header): q8_K per 256 (ggml-quants.c:2555-2592) and q8_0 per 32
// (ggml-cpu/arch/x86/quants.c:297-360).
#pragma once
#include "common.cuh"

namespace b200 {

// Shared-memory copies of the record (stream_decode.cu) store the int8 plane with a 16-byte-segment ROTATION: segment i (0..7) of the r-th
// 128-byte region lives at position (i + r) & 7 of that region.  The register-resident fragment fill of the decode engine (hfrag_fill: lane l
// reads the 8 segments of region l, one per LDS.128) then touches 8 different 16-byte bank groups per quarter-warp instead of one — with the
// previous XOR swizzle that fill was 4-way bank-conflicted and cost ~1 us per matvec phase (measured with the in-kernel marks).
// SWZ = false: linear (global records).
// round-to-nearest-even of |x| < 2^22 exactly as the reference does it (nearest_int, ggml-quants.c: add 1.5 * 2^23, read the mantissa): an FADD
// and an integer subtract on the full-rate pipes instead of F2I on the quarter-rate conversion unit — bit-identical to __float2int_rn here
__device__ __forceinline__ int nearest_int_magic(float x) { return __float_as_int(__fadd_rn(x, 12582912.0f)) - 0x4B400000; }

template <bool SWZ> __device__ __forceinline__ int64_t act_qs_off(int64_t off) {
    if (!SWZ) return off;
    const int64_t seg = off >> 4, r = seg >> 3;
    return (off & 15) | ((r * 8 + ((seg + r) & 7)) << 4);
}

// NB super-blocks in LOCKSTEP (arithmetic of quantize_row_q8_K_ref, ggml-quants.c:2555-2592, bit for bit): the shuffle / divide chains of the blocks are independent, so the warp's
// second block costs almost nothing extra — in the decode engine's prologue 4 of the 12 warps own two blocks and were the critical path.
// live[t] = false: block t is a dummy (nothing is stored).
template <bool SWZ, int NB>
__device__ __forceinline__ void quant_blocks_q8K(const float (&v)[NB][8], const int (&blk)[NB], const bool (&live)[NB], uint8_t * rec, int64_t d_off, int64_t bsum_off) {
    const int lane = threadIdx.x & 31;
    // amax by a plain max tree (one shuffle per step and block); the FIRST element attaining it (the reference keeps the first maximum, whose
    // SIGN enters iscale) = first such element of the lowest lane that holds one (ballot)
    float amax[NB], lmax[NB], first[NB];
#pragma unroll
    for (int t = 0; t < NB; ++t) {
        lmax[t] = 0.0f; first[t] = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float a = fabsf(v[t][i]); if (a > lmax[t]) { lmax[t] = a; first[t] = v[t][i]; } }   // strict '>': first of the lane
        amax[t] = lmax[t];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int t = 0; t < NB; ++t) amax[t] = fmaxf(amax[t], __shfl_xor_sync(0xffffffffu, amax[t], o));
    }
    float vmax[NB];
#pragma unroll
    for (int t = 0; t < NB; ++t) {
        const unsigned holders = __ballot_sync(0xffffffffu, lmax[t] == amax[t]);
        vmax[t] = __shfl_sync(0xffffffffu, first[t], __ffs(holders) - 1);
    }
#pragma unroll
    for (int t = 0; t < NB; ++t) {
        int8_t q[8]; int s = 0;
        float d = 0.0f;
        if (amax[t] != 0.0f) {
            const float iscale = __fdiv_rn(-127.0f, vmax[t]);
#pragma unroll
            for (int i = 0; i < 8; ++i) { int x = nearest_int_magic(__fmul_rn(iscale, v[t][i])); x = x > 127 ? 127 : x; q[i] = (int8_t) x; s += x; }
            d = __fdiv_rn(1.0f, iscale);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = 0;
        }
        const int s2 = s + __shfl_xor_sync(0xffffffffu, s, 1);        // 16-wide partial sums
        if (live[t]) {
            *(uint2 *) (rec + act_qs_off<SWZ>((int64_t) blk[t] * 256 + lane * 8)) = *(const uint2 *) q;
            if ((lane & 1) == 0) ((int16_t *) (rec + bsum_off))[blk[t] * 16 + (lane >> 1)] = (int16_t) s2;
            if (lane == 0) ((float *) (rec + d_off))[blk[t]] = d;
        }
    }
}

// One warp quantises one 256-element super-block held in `v` (lane owns elements [8*lane, 8*lane+8)) into the planar record: the NB = 1 case
// of quant_blocks_q8K, so that the bit-exactness tests of the stand-alone quantiser (tests/test_gpu_parity.py) cover the engine's code too.
template <bool SWZ = false>
__device__ __forceinline__ void quant_block_q8K(const float (&v)[8], uint8_t * rec, int64_t blk, int64_t d_off, int64_t bsum_off) {
    float vv[1][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) vv[0][i] = v[i];
    const int blks[1] = { (int) blk }; const bool lives[1] = { true };
    quant_blocks_q8K<SWZ, 1>(vv, blks, lives, rec, d_off, bsum_off);
}

// 8 lanes quantise one 32-element block (lane part = lane & 7 owns 4 elements); `live` = block index in range.
template <bool SWZ = false>
__device__ __forceinline__ void quant_block_q8_0(const float4 v, bool live, uint8_t * rec, int64_t blk, int64_t d_off, int64_t bsum_off) {
    const int lane = threadIdx.x & 31;
    float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float d  = __fdiv_rn(amax, 127.0f);
    const float id = amax != 0.0f ? __fdiv_rn(127.0f, amax) : 0.0f;
    const int q0 = __float2int_rn(__fmul_rn(v.x, id)), q1 = __float2int_rn(__fmul_rn(v.y, id));
    const int q2 = __float2int_rn(__fmul_rn(v.z, id)), q3 = __float2int_rn(__fmul_rn(v.w, id));
    int s = q0 + q1 + q2 + q3;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (!live) return;
    const uint32_t packed = (uint32_t) (uint8_t) q0 | ((uint32_t) (uint8_t) q1 << 8) | ((uint32_t) (uint8_t) q2 << 16) | ((uint32_t) (uint8_t) q3 << 24);
    *(uint32_t *) (rec + act_qs_off<SWZ>(blk * 32 + (lane & 7) * 4)) = packed;
    if ((lane & 7) == 0) {
        ((__half *) (rec + d_off))[blk] = __float2half_rn(d);
        ((int16_t *) (rec + bsum_off))[blk] = (int16_t) s;
    }
}

// A whole CTA quantises `k` floats at `src` (global or shared, 16-byte aligned) into one record.
__device__ __forceinline__ void quant_row_cta(const float * src, uint8_t * rec, int64_t k, const ActLayout & L) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (L.group == 256) {
        for (int64_t blk = warp; blk < k / 256; blk += nw) {
            const float * xs = src + blk * 256 + lane * 8;
            const float4 v0 = *(const float4 *) xs, v1 = *(const float4 *) (xs + 4);
            const float v[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
            quant_block_q8K(v, rec, blk, L.d_off, L.bsum_off);
        }
    } else {
        const int64_t nblk = k / 32;
        for (int64_t b0 = (int64_t) warp * 4; b0 < nblk; b0 += (int64_t) nw * 4) {
            const int64_t blk = b0 + (lane >> 3);
            const bool live = blk < nblk;
            float4 v = make_float4(0, 0, 0, 0);
            if (live) v = *(const float4 *) (src + blk * 32 + (lane & 7) * 4);
            quant_block_q8_0(v, live, rec, blk, L.d_off, L.bsum_off);
        }
    }
}

} // namespace b200
