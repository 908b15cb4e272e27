// fused_decode.cu — fused producers that take launches off the token critical path (SURVEY.md §7 "hard parts": the reference
// spends 868 node launches + 253 quantize_q8_1 launches per token; the HBM floor is 0.71 ms).
//
//   b200_rms_norm_quantize : RMS_NORM -> MUL(weight) -> activation quantisation in ONE launch.  Replaces rms_norm_f32<.., do_mul>
//                            (ggml-cuda/norm.cu:107-185) followed by quantize_q8_1 (quantize.cu:4-48) before every weight matvec.
//   b200_qkv_post          : per-head RMS_NORM(q), RMS_NORM(k) -> MUL -> ROPE -> K/V cache write (SET_ROWS to F16) in ONE launch.
//                            Replaces 2 x (rms_norm_f32, rope_neox) + 2 x k_set_rows (norm.cu, rope.cu:83-123, set-rows.cu:264) =
//                            6 launches per layer.
// Arithmetic is the same sequence of f32 operations the separate ops perform (ops_misc.cu), so fused and unfused paths agree bit
// for bit except for the order of the sum of squares.
#include "quant_dev.cuh"
#include <math.h>

namespace b200 {

__device__ __forceinline__ float block_sum_f(float v, float * red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = lane < nw ? red[lane] : 0.0f;
    t = warp_sum(t);
    __syncthreads();
    return t;
}

// one CTA per activation column; the normalised row lives in shared memory between the two passes
__global__ void __launch_bounds__(512) k_rms_norm_quantize(const float * __restrict__ x, int64_t x_col_stride, const float * __restrict__ w,
                                                           float * __restrict__ y, int64_t y_col_stride, uint8_t * __restrict__ act,
                                                           int weight_type, int64_t k, float eps) {
    extern __shared__ __align__(16) float row[];
    __shared__ float red[32];
    const int64_t col = blockIdx.x;
    const float * xs = x + col * x_col_stride;
    float ss = 0.0f;
    for (int64_t i = threadIdx.x * 4; i < k; i += blockDim.x * 4) {
        const float4 v = *(const float4 *) (xs + i);
        *(float4 *) (row + i) = v;
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = block_sum_f(ss, red);
    const float scale = 1.0f / sqrtf(ss / (float) k + eps);
    for (int64_t i = threadIdx.x * 4; i < k; i += blockDim.x * 4) {
        float4 v = *(const float4 *) (row + i);
        const float4 ww = *(const float4 *) (w + i);
        v.x = __fmul_rn(__fmul_rn(v.x, scale), ww.x); v.y = __fmul_rn(__fmul_rn(v.y, scale), ww.y);
        v.z = __fmul_rn(__fmul_rn(v.z, scale), ww.z); v.w = __fmul_rn(__fmul_rn(v.w, scale), ww.w);
        *(float4 *) (row + i) = v;
        if (y) *(float4 *) (y + col * y_col_stride + i) = v;
    }
    __syncthreads();
    const ActLayout L = act_layout(weight_type, k);
    quant_row_cta(row, act + col * L.bytes, k, L);
}

// ---- qkv_post -----------------------------------------------------------------------------------------------------------------
struct QkvPostArgs {
    float * q; const float * k; const float * v; const float * qw; const float * kw;
    const int32_t * pos; const void * idx; int idx_i64;
    char * kc; char * vc; int64_t kc_row, vc_row;            // cache row strides in bytes
    int n_head, n_head_kv, mode, has_norm; int64_t q_tok, k_tok, v_tok;   // token strides in elements
    float eps, theta_scale, freq_scale, ext_factor, attn_factor, corr0, corr1;
};

// (cos, sin) * mscale of pair p at position posf: the oracle's theta chain (theta *= theta_scale per pair, ops.cpp ggml_rope_cache_init), YaRN mix, accurate sincosf
__device__ __forceinline__ float2 rope_cs(float posf, int p, float theta_scale, float freq_scale, float ext_factor, float attn_factor, float corr0, float corr1) {
    float theta = posf;
    for (int j = 0; j < p; ++j) theta = __fmul_rn(theta, theta_scale);
    float th = freq_scale * theta, ms = attn_factor;
    if (ext_factor != 0.0f) {
        const float yv = ((float) p - corr0) / fmaxf(0.001f, corr1 - corr0);
        const float ramp = (1.0f - fminf(1.0f, fmaxf(0.0f, yv))) * ext_factor;
        th = __fadd_rn(__fmul_rn(th, 1.0f - ramp), __fmul_rn(theta, ramp));
        ms *= 1.0f + 0.1f * logf(1.0f / freq_scale);
    }
    float sn, cs; sincosf(th, &sn, &cs);
    return make_float2(cs * ms, sn * ms);
}

// CTA per token: the D/2 rotation angles of the token are computed ONCE into shared memory (every head of the token rotates by the same angles: per-element
// recomputation of the 63-step theta chain + sincosf was 3/4 of this kernel's time), then the warps walk the token's heads:
// heads [0, n_head) = Q (in place), [n_head, n_head + n_head_kv) = K -> cache, then V -> cache
template <int D>
__global__ void __launch_bounds__(256) k_qkv_post(const QkvPostArgs A) {
    constexpr int E = D / 32;                                  // elements per lane, contiguous
    __shared__ float2 tab[D / 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t t = blockIdx.x;
    if (threadIdx.x < D / 2) tab[threadIdx.x] = rope_cs((float) A.pos[t], threadIdx.x, A.theta_scale, A.freq_scale, A.ext_factor, A.attn_factor, A.corr0, A.corr1);
    __syncthreads();
    const int nqk = A.n_head + A.n_head_kv;
    const int64_t row = A.idx_i64 ? ((const int64_t *) A.idx)[t] : (int64_t) ((const int32_t *) A.idx)[t];
    for (int h = warp; h < nqk + A.n_head_kv; h += nw) {
    if (h >= nqk) {                                            // V head: convert and store
        const int hv = h - nqk;
        const float * src = A.v + t * A.v_tok + (int64_t) hv * D + lane * E;
        __half * dst = (__half *) (A.vc + row * A.vc_row) + (int64_t) hv * D + lane * E;
#pragma unroll
        for (int i = 0; i < E; ++i) dst[i] = __float2half_rn(src[i]);
        continue;
    }
    const bool is_q = h < A.n_head;
    const int hh = is_q ? h : h - A.n_head;
    const float * src = is_q ? A.q + t * A.q_tok + (int64_t) hh * D : A.k + t * A.k_tok + (int64_t) hh * D;
    float v[E];
#pragma unroll
    for (int i = 0; i < E; ++i) v[i] = src[lane * E + i];
    if (A.has_norm) {
        float ss = 0.0f;
#pragma unroll
        for (int i = 0; i < E; ++i) ss += v[i] * v[i];
        ss = warp_sum(ss);
        const float scale = 1.0f / sqrtf(ss / (float) D + A.eps);
        const float * w = is_q ? A.qw : A.kw;
#pragma unroll
        for (int i = 0; i < E; ++i) v[i] = __fmul_rn(__fmul_rn(v[i], scale), w[lane * E + i]);
    }
    // rotation: neox pairs (p, p + D/2) live in lanes (l, l + 16); norm pairs (2p, 2p + 1) live inside a lane
    float out[E];
    if (A.mode & 2) {
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const float2 cs = tab[(lane & 15) * E + i];
            const float other = __shfl_xor_sync(0xffffffffu, v[i], 16);
            out[i] = lane < 16 ? v[i] * cs.x - other * cs.y : other * cs.y + v[i] * cs.x;
        }
    } else {
#pragma unroll
        for (int i = 0; i < E; i += 2) {
            const float2 cs = tab[(lane * E + i) / 2];
            out[i] = v[i] * cs.x - v[i + 1] * cs.y; out[i + 1] = v[i] * cs.y + v[i + 1] * cs.x;
        }
    }
    if (is_q) {
        float * dst = A.q + t * A.q_tok + (int64_t) hh * D + lane * E;
#pragma unroll
        for (int i = 0; i < E; ++i) dst[i] = out[i];
    } else {
        __half * dst = (__half *) (A.kc + row * A.kc_row) + (int64_t) hh * D + lane * E;
#pragma unroll
        for (int i = 0; i < E; ++i) dst[i] = __float2half_rn(out[i]);
    }
    }
}

static float yarn_corr_dim_f(int n_dims, int n_ctx_orig, float n_rot, float base) {
    return n_dims * logf(n_ctx_orig / (n_rot * 2 * (float) M_PI)) / (2 * logf(base));
}

} // namespace b200

using namespace b200;

extern "C" int b200_rms_norm_quantize(const float * x, int64_t x_col_stride, const float * w, float * y_f32_or_null, int64_t y_col_stride,
                                      void * act, int weight_type, int64_t k, int64_t ncols, float eps, void * stream) {
    if (!x || !w || !act) return B200_ERR_ARG;
    if (!is_quant(weight_type) || k <= 0 || k % blck_size(weight_type) || k % 4 || k > 16384 || ncols <= 0) return B200_ERR_UNSUPPORTED;
    if (((uintptr_t) x | (uintptr_t) w | (uintptr_t) y_f32_or_null | (uintptr_t) act) % 16 || x_col_stride % 4 || y_col_stride % 4) return B200_ERR_UNSUPPORTED;
    static smem_mask_t attr_set{0};
    B200_CUDA_TRY(ensure_dyn_smem(k_rms_norm_quantize, 16384 * 4, attr_set));
    k_rms_norm_quantize<<<(unsigned) ncols, 512, (size_t) k * 4, (cudaStream_t) stream>>>(x, x_col_stride, w, y_f32_or_null, y_col_stride,
                                                                                        (uint8_t *) act, weight_type, k, eps);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_qkv_post(float * q, const float * k, const float * v, const float * q_norm_w, const float * k_norm_w,
                             const int32_t * pos, const void * kv_idx, int idx_type, void * k_cache, void * v_cache,
                             int64_t k_row_stride_bytes, int64_t v_row_stride_bytes, int head_dim, int n_head, int n_head_kv,
                             int64_t n_tok, int64_t q_tok_stride, int64_t k_tok_stride, int64_t v_tok_stride,
                             const b200_rope_params * p, float eps, void * stream) {
    if (!q || !k || !v || !pos || !kv_idx || !k_cache || !v_cache || !p) return B200_ERR_ARG;
    if ((head_dim != 64 && head_dim != 128) || p->n_dims != head_dim || (p->mode != 0 && p->mode != 2)) return B200_ERR_UNSUPPORTED;
    if ((idx_type != B200_I64 && idx_type != B200_I32) || (q_norm_w == nullptr) != (k_norm_w == nullptr)) return B200_ERR_UNSUPPORTED;
    if (n_tok <= 0) return B200_OK;
    if (n_tok > 65535) return B200_ERR_UNSUPPORTED;
    QkvPostArgs A = {};
    A.q = q; A.k = k; A.v = v; A.qw = q_norm_w; A.kw = k_norm_w; A.pos = pos; A.idx = kv_idx; A.idx_i64 = idx_type == B200_I64;
    A.kc = (char *) k_cache; A.vc = (char *) v_cache; A.kc_row = k_row_stride_bytes; A.vc_row = v_row_stride_bytes;
    A.n_head = n_head; A.n_head_kv = n_head_kv; A.mode = p->mode; A.has_norm = q_norm_w != nullptr;
    A.q_tok = q_tok_stride; A.k_tok = k_tok_stride; A.v_tok = v_tok_stride; A.eps = eps;
    A.theta_scale = powf(p->freq_base, -2.0f / p->n_dims);
    A.freq_scale = p->freq_scale; A.ext_factor = p->ext_factor; A.attn_factor = p->attn_factor;
    const float lo = floorf(yarn_corr_dim_f(p->n_dims, p->n_ctx_orig, p->beta_fast, p->freq_base));
    const float hi = ceilf (yarn_corr_dim_f(p->n_dims, p->n_ctx_orig, p->beta_slow, p->freq_base));
    A.corr0 = lo < 0 ? 0 : lo; A.corr1 = hi > p->n_dims - 1 ? p->n_dims - 1 : hi;
    if (head_dim == 128) k_qkv_post<128><<<(unsigned) n_tok, 256, 0, (cudaStream_t) stream>>>(A);
    else                 k_qkv_post<64><<<(unsigned) n_tok, 256, 0, (cudaStream_t) stream>>>(A);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
