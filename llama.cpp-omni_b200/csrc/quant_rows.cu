// quant_rows.cu — rows in and out of the ggml block formats: the data formats ADJACENT to the hot path.
//   F32 rows  -> q8_0 / q4_0 blocks   SET_ROWS into a quantised KV cache (-ctk / -ctv q8_0 | q4_0) and CPY F32 -> quant (K-shift of a quantised cache)
//                                     replaces k_set_rows_quant (ggml-cuda/set-rows.cu:8,75) and cpy_f32_q (cpy.cu); CPU: ggml_compute_forward_set_rows_f32 ->
//                                     type_traits_cpu.from_float = quantize_row_q8_0 (x86 build: ggml-cpu/arch/x86/quants.c:297-360, id = 127/amax, nearest-even)
//                                     / quantize_row_q4_0 -> quantize_row_q4_0_ref (ggml-quants.c:36-71)
//   quant rows -> F32 / F16           GET_ROWS on a quantised token_embd (native or planar weights), CPY quant -> F32, and the F16 staging of a quantised K / V
//                                     for FLASH_ATTN_EXT; replaces k_get_rows (getrows.cu:6), dequantize_block (convert.cu) and the to_fp16 staging of
//                                     launch_fattn (fattn-common.cuh); CPU: dequantize_row_* (ggml-quants.c:307-325, 390-402, 1352-1374, 1554-1578, 1762-1791)
// HBM-bound byte shuffling: coalesced 16-byte loads of the F32 side (8 lanes per 32-element block), one element per thread on the dequantising side.
#include "common.cuh"

namespace b200 {

struct R4 { char * data; int type; int64_t ne[4]; int64_t nb[4]; };
static inline R4 r4(const b200_tensor * t) {
    R4 r; r.data = (char *) t->data; r.type = t->type;
    for (int i = 0; i < 4; ++i) { r.ne[i] = t->ne[i]; r.nb[i] = t->nb[i]; }
    return r;
}
static inline unsigned qgrid(int64_t threads) {
    int64_t g = (threads + 255) / 256;
    const int64_t cap = (int64_t) sm_count() * 16;
    return (unsigned) (g > cap ? cap : g < 1 ? 1 : g);
}

// ---------------------------------------------------------------------------------------------------------------- F32 -> q8_0 / q4_0
struct QrArgs { R4 src, idx, dst; int has_idx; };

// 8 lanes own one 32-element block (4 consecutive floats each); writes the NATIVE block (f16 d + payload, 2-byte aligned)
template <int T>
__global__ void __launch_bounds__(256) k_quant_rows(const QrArgs A, int64_t nblk_total) {
    const int lane = threadIdx.x & 31, part = lane & 7;
    const int64_t nb_row = A.src.ne[0] / 32;
    const int64_t ngroups = (nblk_total + 3) / 4 * 4;                         // whole warps: shuffles need every lane
    for (int64_t gb = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 3; gb < ngroups; gb += ((int64_t) gridDim.x * blockDim.x) >> 3) {
        const bool live = gb < nblk_total;
        const int64_t b = live ? gb % nb_row : 0, row = live ? gb / nb_row : 0;
        const int64_t i1 = row % A.src.ne[1], i2 = (row / A.src.ne[1]) % A.src.ne[2], i3 = row / (A.src.ne[1] * A.src.ne[2]);
        float4 v = make_float4(0, 0, 0, 0);
        if (live) {
            const char * sp = A.src.data + i1 * A.src.nb[1] + i2 * A.src.nb[2] + i3 * A.src.nb[3] + (b * 32 + part * 4) * 4;
            if (((uintptr_t) sp & 15) == 0) v = *(const float4 *) sp;
            else { const float * f = (const float *) sp; v = make_float4(f[0], f[1], f[2], f[3]); }
        }
        int64_t r = i1;
        if (A.has_idx && live) {
            const char * ip = A.idx.data + i1 * A.idx.nb[0] + (i2 % A.idx.ne[1]) * A.idx.nb[1] + (i3 % A.idx.ne[2]) * A.idx.nb[2];
            r = A.idx.type == B200_I64 ? *(const int64_t *) ip : (int64_t) *(const int32_t *) ip;
        }
        uint8_t * blk = (uint8_t *) A.dst.data + r * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3] + b * (T == B200_Q8_0 ? 34 : 18);
        if (T == B200_Q8_0) {
            float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            const float d  = __fdiv_rn(amax, 127.0f);
            const float id = amax != 0.0f ? __fdiv_rn(127.0f, amax) : 0.0f;
            const int q0 = __float2int_rn(__fmul_rn(v.x, id)), q1 = __float2int_rn(__fmul_rn(v.y, id));
            const int q2 = __float2int_rn(__fmul_rn(v.z, id)), q3 = __float2int_rn(__fmul_rn(v.w, id));
            if (live) {
                uint16_t * o16 = (uint16_t *) (blk + 2 + part * 4);
                o16[0] = (uint16_t) ((uint8_t) q0 | ((uint32_t) (uint8_t) q1 << 8));
                o16[1] = (uint16_t) ((uint8_t) q2 | ((uint32_t) (uint8_t) q3 << 8));
                if (part == 0) *(__half *) blk = __float2half_rn(d);
            }
        } else {
            // the element of largest magnitude, FIRST one on ties (the reference scans with a strict `<`), keeps its sign in d = max / -8
            const float xs[4] = { v.x, v.y, v.z, v.w };
            float amax = 0.0f, vmax = 0.0f; int imax = 32;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (amax < fabsf(xs[j])) { amax = fabsf(xs[j]); vmax = xs[j]; imax = part * 4 + j; }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                const float a2 = __shfl_xor_sync(0xffffffffu, amax, o), v2 = __shfl_xor_sync(0xffffffffu, vmax, o);
                const int   i2_ = __shfl_xor_sync(0xffffffffu, imax, o);
                if (a2 > amax || (a2 == amax && i2_ < imax)) { amax = a2; vmax = v2; imax = i2_; }
            }
            const float d  = __fdiv_rn(vmax, -8.0f);
            const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
            uint32_t code = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int q = (int) (int8_t) __float2int_rz(__fadd_rn(__fmul_rn(xs[j], id), 8.5f));
                q = q > 15 ? 15 : q;
                code |= (uint32_t) (q & 0xff) << (8 * j);
            }
            const uint32_t other = __shfl_xor_sync(0xffffffffu, code, 4);       // elements j + 16 of the same block (lanes part + 4)
            if (live && part < 4) {
                const uint32_t packed = (code & 0x0f0f0f0fu) | ((other & 0x0f0f0f0fu) << 4);
                uint16_t * o16 = (uint16_t *) (blk + 2 + part * 4);
                o16[0] = (uint16_t) packed; o16[1] = (uint16_t) (packed >> 16);
                if (part == 0) *(__half *) blk = __float2half_rn(d);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------- quant -> F32 / F16
__device__ __forceinline__ void scale_min_k4(const uint8_t * p, int j, int & sc, int & mn) {                  // ggml-quants.c:703-711
    if (j < 4) { sc = p[j] & 63; mn = p[j + 4] & 63; }
    else       { sc = (p[j + 4] & 15) | ((p[j - 4] >> 6) << 4); mn = (p[j + 4] >> 4) | ((p[j] >> 6) << 4); }
}
// element e of one block; `pay` = payload (q4_0 / q8_0: the 16 / 32 qs bytes; q6_K: ql | qh | scales; q4_K / q5_K: the whole block), dbits = the block's f16 d
__device__ __forceinline__ float deq_elem(int type, const uint8_t * pay, uint16_t dbits, int e) {
    switch (type) {
        case B200_Q4_0: { const int c = (e < 16 ? pay[e] & 15 : pay[e - 16] >> 4) - 8; return __fmul_rn((float) c, h2f(dbits)); }
        case B200_Q8_0: return __fmul_rn((float) (int8_t) pay[e], h2f(dbits));
        case B200_Q4_K: {
            int sc, mn; scale_min_k4(pay + 4, e >> 5, sc, mn);
            const float d = __fmul_rn(h2f(*(const uint16_t *) pay), (float) sc), m = __fmul_rn(h2f(*(const uint16_t *) (pay + 2)), (float) mn);
            const uint8_t q = pay[16 + 32 * (e >> 6) + (e & 31)];
            return d * (float) ((e & 32) ? q >> 4 : q & 15) - m;
        }
        case B200_Q5_K: {
            int sc, mn; scale_min_k4(pay + 4, e >> 5, sc, mn);
            const float d = __fmul_rn(h2f(*(const uint16_t *) pay), (float) sc), m = __fmul_rn(h2f(*(const uint16_t *) (pay + 2)), (float) mn);
            const uint8_t q = pay[48 + 32 * (e >> 6) + (e & 31)];
            const int hi = (pay[16 + (e & 31)] >> (e >> 5)) & 1;
            return d * (float) (((e & 32) ? q >> 4 : q & 15) + 16 * hi) - m;
        }
        case B200_Q6_K: {
            const int h = e >> 7, r = e & 127, l = r & 31, quad = r >> 5;
            int lo = (quad & 1) ? pay[64 * h + 32 + l] : pay[64 * h + l];
            lo = (quad & 2) ? lo >> 4 : lo & 15;
            const int c = (lo | (((pay[128 + 32 * h + l] >> (2 * quad)) & 3) << 4)) - 32;
            return __fmul_rn(__fmul_rn(h2f(dbits), (float) (int8_t) pay[192 + (e >> 4)]), (float) c);
        }
    }
    return 0.0f;
}

struct DqArgs { R4 src, idx, dst; int has_idx, planar; int64_t nblocks_total; };
// dst [ne0, n, ne2, ne3] <- row r of src (r = idx[i1, i2, i3] for GET_ROWS, i1 otherwise)
template <typename TD>
__global__ void __launch_bounds__(256) k_dequant_rows(const DqArgs A, int64_t total) {
    const int type = A.src.type, qk = blck_size(type), ts = type_size(type), ps = payload_size(type);
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t) gridDim.x * blockDim.x) {
        int64_t t = g;
        const int64_t i0 = t % A.dst.ne[0]; t /= A.dst.ne[0];
        const int64_t i1 = t % A.dst.ne[1]; t /= A.dst.ne[1];
        const int64_t i2 = t % A.dst.ne[2], i3 = t / A.dst.ne[2];
        int64_t r = i1;
        if (A.has_idx) r = (int64_t) *(const int32_t *) (A.idx.data + i1 * A.idx.nb[0] + i2 * A.idx.nb[1] + i3 * A.idx.nb[2]);
        const int64_t b = i0 / qk; const int e = (int) (i0 % qk);
        const uint8_t * pay; uint16_t dbits = 0;
        if (A.planar) {                                                      // contiguous weight tensor: payload plane, then the f16 d plane
            const int64_t nb_row = A.src.ne[0] / qk;
            const int64_t bi = ((i3 * A.src.ne[2] + i2) * A.src.ne[1] + r) * nb_row + b;
            pay = (const uint8_t *) A.src.data + bi * ps;
            dbits = *(const uint16_t *) ((const uint8_t *) A.src.data + A.nblocks_total * ps + bi * 2);
        } else {
            const uint8_t * blk = (const uint8_t *) A.src.data + r * A.src.nb[1] + (i2 % A.src.ne[2]) * A.src.nb[2] + (i3 % A.src.ne[3]) * A.src.nb[3] + b * ts;
            if (type == B200_Q4_0 || type == B200_Q8_0) { dbits = *(const uint16_t *) blk; pay = blk + 2; }
            else if (type == B200_Q6_K)                 { dbits = *(const uint16_t *) (blk + 208); pay = blk; }
            else pay = blk;
        }
        const float v = deq_elem(type, pay, dbits, e);
        char * dp = A.dst.data + i0 * A.dst.nb[0] + i1 * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3];
        if (sizeof(TD) == 4) *(float *) dp = v; else *(__half *) dp = __float2half_rn(v);
    }
}

// entry points used by ops_misc.cu (SET_ROWS / GET_ROWS / CPY) and flash_attn.cu (F16 staging of a quantised K / V)
int quant_rows(const b200_tensor * src, const b200_tensor * idx, const b200_tensor * dst, cudaStream_t st) {
    if (src->type != B200_F32 || (dst->type != B200_Q8_0 && dst->type != B200_Q4_0) || src->nb[0] != 4 || src->ne[0] % 32 || dst->ne[0] != src->ne[0]) return B200_ERR_UNSUPPORTED;
    if (((uintptr_t) dst->data | dst->nb[1] | dst->nb[2] | dst->nb[3]) & 1) return B200_ERR_UNSUPPORTED;
    const int64_t nblk = src->ne[0] / 32 * src->ne[1] * src->ne[2] * src->ne[3];
    if (nblk == 0) return B200_OK;
    QrArgs A; A.src = r4(src); A.dst = r4(dst); A.has_idx = idx != nullptr; if (idx) A.idx = r4(idx); else A.idx = A.src;
    if (dst->type == B200_Q8_0) k_quant_rows<B200_Q8_0><<<qgrid(nblk * 8), 256, 0, st>>>(A, nblk);
    else                        k_quant_rows<B200_Q4_0><<<qgrid(nblk * 8), 256, 0, st>>>(A, nblk);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int dequant_rows(const b200_tensor * src, const b200_tensor * idx, const b200_tensor * dst, cudaStream_t st) {
    if (!is_quant(src->type) || (dst->type != B200_F32 && dst->type != B200_F16) || src->ne[0] != dst->ne[0] || src->ne[0] % blck_size(src->type)) return B200_ERR_UNSUPPORTED;
    if (idx && idx->type != B200_I32) return B200_ERR_UNSUPPORTED;
    const bool planar = src->layout == B200_LAYOUT_PLANAR && payload_size(src->type) != type_size(src->type);
    const int64_t total = dst->ne[0] * dst->ne[1] * dst->ne[2] * dst->ne[3];
    if (total == 0) return B200_OK;
    DqArgs A; A.src = r4(src); A.dst = r4(dst); A.has_idx = idx != nullptr; if (idx) A.idx = r4(idx); else A.idx = A.src;
    A.planar = planar; A.nblocks_total = src->ne[0] / blck_size(src->type) * src->ne[1] * src->ne[2] * src->ne[3];
    if (planar) {                                                            // planes are addressed by block index: the tensor must be whole and contiguous
        const int64_t rb = src->ne[0] / blck_size(src->type) * type_size(src->type);
        if (src->nb[1] != rb || src->nb[2] != rb * src->ne[1] || src->nb[3] != rb * src->ne[1] * src->ne[2]) return B200_ERR_UNSUPPORTED;
    }
    if (dst->type == B200_F32) k_dequant_rows<float><<<qgrid(total), 256, 0, st>>>(A, total);
    else                       k_dequant_rows<__half><<<qgrid(total), 256, 0, st>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

} // namespace b200
