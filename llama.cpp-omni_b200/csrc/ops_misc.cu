// ops_misc.cu — the non-matmul ops of the decode/prefill graph, one sm_100a kernel each:
//   RMS_NORM(+MUL+ADD)  replaces rms_norm_f32<block, do_mul, do_add>   ggml-cuda/norm.cu:107-185     oracle ops.cpp:3517-3565
//   ROPE norm/neox      replaces rope_norm / rope_neox                   ggml-cuda/rope.cu:40-123      oracle ops.cpp:5436-5720
//   SET_ROWS            replaces k_set_rows                              ggml-cuda/set-rows.cu:264
//   GET_ROWS            replaces k_get_rows_float                        ggml-cuda/getrows.cu:238
//   CPY/CONT/DUP        replaces cpy_flt                                 ggml-cuda/cpy.cu:280
//   ADD/SUB/MUL/DIV     replaces k_bin_bcast                             ggml-cuda/binbcast.cu:395-443
//   unary / GLU / SCALE replaces unary_op_kernel, unary_gated_op_kernel  ggml-cuda/unary.cu:208-292, scale.cu
//   SOFT_MAX            replaces soft_max_f32                            ggml-cuda/softmax.cu:253
// All of them move < 1 % of the decode bytes; what matters is that each is ONE launch with 128-bit accesses where the layout
// allows and that the fused variants (rms_norm->mul->quantise, see fused_decode.cu) remove launches from the token critical path.
#include "common.cuh"
#include <math.h>

namespace b200 {

struct T4 {                    // device-side copy of b200_tensor
    char * data; int type; int64_t ne[4]; int64_t nb[4];
};
static inline T4 t4(const b200_tensor * t) {
    T4 r; r.data = (char *) t->data; r.type = t->type;
    for (int i = 0; i < 4; ++i) { r.ne[i] = t->ne[i]; r.nb[i] = t->nb[i]; }
    return r;
}
static inline int64_t nelem(const b200_tensor * t) { return t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3]; }
static inline int64_t nrows(const b200_tensor * t) { return t->ne[1] * t->ne[2] * t->ne[3]; }
static inline bool same_shape(const b200_tensor * a, const b200_tensor * b) {
    return a->ne[0] == b->ne[0] && a->ne[1] == b->ne[1] && a->ne[2] == b->ne[2] && a->ne[3] == b->ne[3];
}
static inline bool is_contig(const b200_tensor * t) {
    const int64_t ts = type_size(t->type);
    return t->nb[0] == ts && t->nb[1] == ts * t->ne[0] && t->nb[2] == t->nb[1] * t->ne[1] && t->nb[3] == t->nb[2] * t->ne[2];
}
static inline unsigned grid_for(int64_t work_items, int per_block) {
    int64_t g = (work_items + per_block - 1) / per_block;
    const int64_t cap = (int64_t) sm_count() * 16;                    // grid-stride beyond 16 CTAs per SM
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned) g;
}

template <typename T> __device__ __forceinline__ float ldf(const void * p);
template <> __device__ __forceinline__ float ldf<float>(const void * p) { return *(const float *) p; }
template <> __device__ __forceinline__ float ldf<__half>(const void * p) { return __half2float(*(const __half *) p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const void * p) { return __bfloat162float(*(const __nv_bfloat16 *) p); }
template <typename T> __device__ __forceinline__ void stf(void * p, float v);
template <> __device__ __forceinline__ void stf<float>(void * p, float v) { *(float *) p = v; }
template <> __device__ __forceinline__ void stf<__half>(void * p, float v) { *(__half *) p = __float2half_rn(v); }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(void * p, float v) { *(__nv_bfloat16 *) p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float ld_any(const char * p, int type) {
    return type == B200_F32 ? ldf<float>(p) : type == B200_F16 ? ldf<__half>(p) : ldf<__nv_bfloat16>(p);
}
__device__ __forceinline__ void st_any(char * p, int type, float v) {
    if (type == B200_F32) stf<float>(p, v); else if (type == B200_F16) stf<__half>(p, v); else stf<__nv_bfloat16>(p, v);
}

// ================================================================== RMS_NORM (+ MUL + ADD) ====================================
// One row per warp (ne0 <= 1024) or per CTA.  Sum of squares in f32 with a fixed tree order (the oracle sums in f64; the
// difference is far below the 1e-7 NMSE the reference's own backend test allows).
struct NormArgs { T4 x, w, add, dst; float eps; int has_w, has_add; int64_t rows; uint8_t * tiles; int write_f32; };     // tiles: optional F16 activation tiles (act_tile_off)

__device__ __forceinline__ float block_sum(float v, float * smem) {      // smem: 32 floats
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    float t = lane < nw ? smem[lane] : 0.0f;
    t = warp_sum(t);
    __syncthreads();
    return t;
}

template <bool WARP_ROWS>
__global__ void __launch_bounds__(256) k_rms_norm(const NormArgs A) {
    __shared__ float red[32];
    const int lane = threadIdx.x & 31;
    const int64_t row = WARP_ROWS ? (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) : blockIdx.x;
    if (WARP_ROWS && row >= A.rows) return;
    const int64_t i1 = row % A.x.ne[1], i2 = (row / A.x.ne[1]) % A.x.ne[2], i3 = row / (A.x.ne[1] * A.x.ne[2]);
    const float * x = (const float *) (A.x.data + i1 * A.x.nb[1] + i2 * A.x.nb[2] + i3 * A.x.nb[3]);
    float * y = (float *) (A.dst.data + i1 * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3]);
    const float * w = A.has_w ? (const float *) (A.w.data + (i1 % A.w.ne[1]) * A.w.nb[1] + (i2 % A.w.ne[2]) * A.w.nb[2] + (i3 % A.w.ne[3]) * A.w.nb[3]) : nullptr;
    const float * ad = A.has_add ? (const float *) (A.add.data + (i1 % A.add.ne[1]) * A.add.nb[1] + (i2 % A.add.ne[2]) * A.add.nb[2] + (i3 % A.add.ne[3]) * A.add.nb[3]) : nullptr;
    const int64_t n = A.x.ne[0];
    const int tid = WARP_ROWS ? lane : threadIdx.x, nt = WARP_ROWS ? 32 : blockDim.x;
    const bool v4 = (n % 4 == 0) && (((uintptr_t) x | (uintptr_t) y) % 16 == 0);
    const int64_t wn = A.has_w ? A.w.ne[0] : 1, an = A.has_add ? A.add.ne[0] : 1;
    // block-per-row fast path (rows of up to 4096: every prefill norm): the row is read ONCE, 4 x 16 bytes per thread stay in registers between the sum of
    // squares and the scaling (the generic path below re-reads it: 3 passes over memory instead of 2)
    if (!WARP_ROWS && v4 && n <= 16 * 256 && (!w || (wn == n && (uintptr_t) w % 16 == 0)) && (!ad || (an == n && (uintptr_t) ad % 16 == 0))) {
        float4 r[4];
        float s2 = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int64_t i = ((int64_t) c * 256 + tid) * 4;
            r[c] = i < n ? *(const float4 *) (x + i) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            s2 += r[c].x * r[c].x + r[c].y * r[c].y + r[c].z * r[c].z + r[c].w * r[c].w;
        }
        s2 = block_sum(s2, red);
        const float sc = 1.0f / sqrtf(s2 / (float) n + A.eps);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int64_t i = ((int64_t) c * 256 + tid) * 4;
            if (i >= n) continue;
            float4 v = make_float4(__fmul_rn(r[c].x, sc), __fmul_rn(r[c].y, sc), __fmul_rn(r[c].z, sc), __fmul_rn(r[c].w, sc));
            if (w)  { const float4 q = *(const float4 *) (w + i);  v = make_float4(__fmul_rn(v.x, q.x), __fmul_rn(v.y, q.y), __fmul_rn(v.z, q.z), __fmul_rn(v.w, q.w)); }
            if (ad) { const float4 q = *(const float4 *) (ad + i); v = make_float4(__fadd_rn(v.x, q.x), __fadd_rn(v.y, q.y), __fadd_rn(v.z, q.z), __fadd_rn(v.w, q.w)); }
            if (A.write_f32) *(float4 *) (y + i) = v;
            if (A.tiles) {                                                   // half of a 16-byte core-matrix row: the MUL_MAT that follows needs no conversion pass
                const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
                *(uint2 *) (A.tiles + act_tile_off(row, i & ~(int64_t) 7, n) + (i & 4) * 2) = make_uint2(*(const uint32_t *) &h0, *(const uint32_t *) &h1);
            }
        }
        return;
    }
    float ss = 0.0f;
    if (v4) for (int64_t i = tid * 4; i < n; i += nt * 4) { const float4 v = *(const float4 *) (x + i); ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
    else    for (int64_t i = tid; i < n; i += nt) { const float v = x[i]; ss += v * v; }
    ss = WARP_ROWS ? warp_sum(ss) : block_sum(ss, red);
    const float scale = 1.0f / sqrtf(ss / (float) n + A.eps);
    for (int64_t i = tid; i < n; i += nt) {
        float v = __fmul_rn(x[i], scale);                                   // separate roundings, like the three CPU ops
        if (w)  v = __fmul_rn(v, w[wn == n ? i : i % wn]);
        if (ad) v = __fadd_rn(v, ad[an == n ? i : i % an]);
        y[i] = v;
    }
}

// ================================================================== ROPE =======================================================
struct RopeArgs {
    T4 x, dst; const int32_t * pos; const float * ff;
    int n_dims, mode; float theta_scale, freq_scale, ext_factor, attn_factor, corr0, corr1;
};
// thread = one rotation pair of one (head, token, batch) row.  theta follows the oracle's chain (theta *= theta_scale per pair,
// ops.cpp ggml_rope_cache_init) so that the angle is bit-identical to the CPU backend's; sincosf is the accurate variant.
template <typename T>
__global__ void __launch_bounds__(256) k_rope(const RopeArgs A, int64_t total_pairs) {
    const int64_t half_n = A.x.ne[0] / 2;                                  // pairs per row incl. pass-through region
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < total_pairs; g += (int64_t) gridDim.x * blockDim.x) {
        const int64_t p = g % half_n, row = g / half_n;
        const int64_t i1 = row % A.x.ne[1], i2 = (row / A.x.ne[1]) % A.x.ne[2], i3 = row / (A.x.ne[1] * A.x.ne[2]);
        const T * x = (const T *) (A.x.data + i1 * A.x.nb[1] + i2 * A.x.nb[2] + i3 * A.x.nb[3]);
        T * y = (T *) (A.dst.data + i1 * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3]);
        if (2 * p >= A.n_dims) {                                            // beyond the rotated dims: copy through
            const int64_t i = A.n_dims + 2 * (p - A.n_dims / 2);
            y[i] = x[i]; y[i + 1] = x[i + 1];
            continue;
        }
        float theta = (float) A.pos[i2];
        for (int j = 0; j < (int) p; ++j) theta = __fmul_rn(theta, A.theta_scale);
        const float ffv = A.ff ? A.ff[p] : 1.0f;
        const float th_extrap = theta / ffv;
        float th = __fmul_rn(A.freq_scale, th_extrap), mscale = A.attn_factor;
        if (A.ext_factor != 0.0f) {
            const float yv = ((float) p - A.corr0) / fmaxf(0.001f, A.corr1 - A.corr0);
            const float ramp = (1.0f - fminf(1.0f, fmaxf(0.0f, yv))) * A.ext_factor;
            th = __fadd_rn(__fmul_rn(th, 1.0f - ramp), __fmul_rn(th_extrap, ramp));     // no FMA contraction: the oracle's rounding
            mscale *= 1.0f + 0.1f * logf(1.0f / A.freq_scale);
        }
        float sn, cs; sincosf(th, &sn, &cs);
        cs *= mscale; sn *= mscale;
        const int64_t a = (A.mode & 2) ? p : 2 * p, b = (A.mode & 2) ? p + A.n_dims / 2 : 2 * p + 1;
        const float x0 = ldf<T>(&x[a]), x1 = ldf<T>(&x[b]);
        stf<T>(&y[a], x0 * cs - x1 * sn);
        stf<T>(&y[b], x0 * sn + x1 * cs);
    }
}

// CTA = one (token, batch) slice: every head of the token rotates by the same n_dims/2 angles, so they are computed ONCE per CTA into shared memory
// (the per-pair theta chain + sincosf was most of k_rope's time at prefill sizes) and the threads then stream the slice's heads.  Same arithmetic.
constexpr int ROPE_TAB = 256;                                                // rotation pairs held in shared memory
template <typename T>
__global__ void __launch_bounds__(256) k_rope_rows(const RopeArgs A) {
    __shared__ float2 tab[ROPE_TAB];
    const int64_t i2 = blockIdx.x, i3 = blockIdx.y;
    const int np = A.n_dims / 2;
    for (int p = threadIdx.x; p < np; p += blockDim.x) {
        float theta = (float) A.pos[i2];
        for (int j = 0; j < p; ++j) theta = __fmul_rn(theta, A.theta_scale);
        const float ffv = A.ff ? A.ff[p] : 1.0f;
        const float th_extrap = theta / ffv;
        float th = __fmul_rn(A.freq_scale, th_extrap), mscale = A.attn_factor;
        if (A.ext_factor != 0.0f) {
            const float yv = ((float) p - A.corr0) / fmaxf(0.001f, A.corr1 - A.corr0);
            const float ramp = (1.0f - fminf(1.0f, fmaxf(0.0f, yv))) * A.ext_factor;
            th = __fadd_rn(__fmul_rn(th, 1.0f - ramp), __fmul_rn(th_extrap, ramp));
            mscale *= 1.0f + 0.1f * logf(1.0f / A.freq_scale);
        }
        float sn, cs; sincosf(th, &sn, &cs);
        tab[p] = make_float2(cs * mscale, sn * mscale);
    }
    __syncthreads();
    const int64_t half_n = A.x.ne[0] / 2, per_slice = half_n * A.x.ne[1];
    for (int64_t g = threadIdx.x; g < per_slice; g += blockDim.x) {
        const int64_t p = g % half_n, i1 = g / half_n;
        const T * x = (const T *) (A.x.data + i1 * A.x.nb[1] + i2 * A.x.nb[2] + i3 * A.x.nb[3]);
        T * y = (T *) (A.dst.data + i1 * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3]);
        if (2 * p >= A.n_dims) { const int64_t i = A.n_dims + 2 * (p - np); y[i] = x[i]; y[i + 1] = x[i + 1]; continue; }
        const float2 cs = tab[p];
        const int64_t a = (A.mode & 2) ? p : 2 * p, b = (A.mode & 2) ? p + np : 2 * p + 1;
        const float x0 = ldf<T>(&x[a]), x1 = ldf<T>(&x[b]);
        stf<T>(&y[a], x0 * cs.x - x1 * cs.y);
        stf<T>(&y[b], x0 * cs.y + x1 * cs.x);
    }
}

// ================================================================== SET_ROWS / GET_ROWS / CPY ==================================
struct RowsArgs { T4 src, idx, dst; };

__global__ void __launch_bounds__(256) k_set_rows(const RowsArgs A, int64_t total) {
    const int64_t ne0 = A.src.ne[0];
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t) gridDim.x * blockDim.x) {
        const int64_t i0 = g % ne0, row = g / ne0;
        const int64_t i1 = row % A.src.ne[1], i2 = (row / A.src.ne[1]) % A.src.ne[2], i3 = row / (A.src.ne[1] * A.src.ne[2]);
        const char * ip = A.idx.data + i1 * A.idx.nb[0] + (i2 % A.idx.ne[1]) * A.idx.nb[1] + (i3 % A.idx.ne[2]) * A.idx.nb[2];
        const int64_t r = A.idx.type == B200_I64 ? *(const int64_t *) ip : (int64_t) *(const int32_t *) ip;
        const float v = *(const float *) (A.src.data + i0 * A.src.nb[0] + i1 * A.src.nb[1] + i2 * A.src.nb[2] + i3 * A.src.nb[3]);
        st_any(A.dst.data + i0 * A.dst.nb[0] + r * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3], A.dst.type, v);
    }
}

__global__ void __launch_bounds__(256) k_get_rows(const RowsArgs A, int64_t total) {
    const int64_t ne0 = A.dst.ne[0];
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t) gridDim.x * blockDim.x) {
        const int64_t i0 = g % ne0, row = g / ne0;
        const int64_t i10 = row % A.dst.ne[1], i11 = (row / A.dst.ne[1]) % A.dst.ne[2], i12 = row / (A.dst.ne[1] * A.dst.ne[2]);
        const int64_t r = (int64_t) *(const int32_t *) (A.idx.data + i10 * A.idx.nb[0] + i11 * A.idx.nb[1] + i12 * A.idx.nb[2]);
        const float v = ld_any(A.src.data + i0 * A.src.nb[0] + r * A.src.nb[1] + i11 * A.src.nb[2] + i12 * A.src.nb[3], A.src.type);
        st_any(A.dst.data + i0 * A.dst.nb[0] + i10 * A.dst.nb[1] + i11 * A.dst.nb[2] + i12 * A.dst.nb[3], A.dst.type, v);
    }
}

struct CpyArgs { T4 src, dst; };
// element g of the flattened index space: src coordinates from src.ne, dst coordinates from dst.ne (ggml CPY allows reshape)
__global__ void __launch_bounds__(256) k_cpy(const CpyArgs A, int64_t total) {
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t) gridDim.x * blockDim.x) {
        int64_t r = g;
        const int64_t s0 = r % A.src.ne[0]; r /= A.src.ne[0];
        const int64_t s1 = r % A.src.ne[1]; r /= A.src.ne[1];
        const int64_t s2 = r % A.src.ne[2]; const int64_t s3 = r / A.src.ne[2];
        r = g;
        const int64_t d0 = r % A.dst.ne[0]; r /= A.dst.ne[0];
        const int64_t d1 = r % A.dst.ne[1]; r /= A.dst.ne[1];
        const int64_t d2 = r % A.dst.ne[2]; const int64_t d3 = r / A.dst.ne[2];
        const char * sp = A.src.data + s0 * A.src.nb[0] + s1 * A.src.nb[1] + s2 * A.src.nb[2] + s3 * A.src.nb[3];
        char * dp = A.dst.data + d0 * A.dst.nb[0] + d1 * A.dst.nb[1] + d2 * A.dst.nb[2] + d3 * A.dst.nb[3];
        if (A.src.type == A.dst.type && (A.src.type == B200_I32)) *(int32_t *) dp = *(const int32_t *) sp;
        else st_any(dp, A.dst.type, ld_any(sp, A.src.type));
    }
}
// contiguous same-type or f32->f16 fast path, 16 bytes of source per thread
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) k_cpy_contig(const TS * __restrict__ s, TD * __restrict__ d, int64_t n) {
    constexpr int V = 16 / sizeof(TS);
    const int64_t nv = n / V;
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < nv; g += (int64_t) gridDim.x * blockDim.x) {
        const uint4 raw = *(const uint4 *) (s + g * V);
        const TS * sv = (const TS *) &raw;
        TD out[V];
#pragma unroll
        for (int i = 0; i < V; ++i) { float f = ldf<TS>(&sv[i]); stf<TD>(&out[i], f); }
#pragma unroll
        for (int i = 0; i < V * (int) sizeof(TD) / 8; ++i) ((uint2 *) (d + g * V))[i] = ((const uint2 *) out)[i];
    }
    for (int64_t g = nv * V + (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (int64_t) gridDim.x * blockDim.x)
        stf<TD>(&d[g], ldf<TS>(&s[g]));
}

// ================================================================== binary / unary / GLU / SCALE ===============================
struct BinArgs { T4 a, b, dst; int op; };
__device__ __forceinline__ float binop(int op, float x, float y) {
    return op == B200_ADD ? x + y : op == B200_SUB ? x - y : op == B200_MUL ? x * y : x / y;
}
__global__ void __launch_bounds__(256) k_binary(const BinArgs A, int64_t total) {
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t) gridDim.x * blockDim.x) {
        int64_t r = g;
        const int64_t i0 = r % A.dst.ne[0]; r /= A.dst.ne[0];
        const int64_t i1 = r % A.dst.ne[1]; r /= A.dst.ne[1];
        const int64_t i2 = r % A.dst.ne[2]; const int64_t i3 = r / A.dst.ne[2];
        const float x = ld_any(A.a.data + i0 * A.a.nb[0] + i1 * A.a.nb[1] + i2 * A.a.nb[2] + i3 * A.a.nb[3], A.a.type);
        const float y = ld_any(A.b.data + (i0 % A.b.ne[0]) * A.b.nb[0] + (i1 % A.b.ne[1]) * A.b.nb[1] + (i2 % A.b.ne[2]) * A.b.nb[2]
                               + (i3 % A.b.ne[3]) * A.b.nb[3], A.b.type);
        st_any(A.dst.data + i0 * A.dst.nb[0] + i1 * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3], A.dst.type, binop(A.op, x, y));
    }
}
// all-F32 contiguous, b either same shape or one row broadcast over rows: float4 path
__global__ void __launch_bounds__(256) k_binary_f32x4(const float * __restrict__ a, const float * __restrict__ b, float * __restrict__ d,
                                                      int64_t n4, int64_t b_n4, int op) {
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (int64_t) gridDim.x * blockDim.x) {
        const float4 x = ((const float4 *) a)[g], y = ((const float4 *) b)[b_n4 == n4 ? g : g % b_n4];
        float4 r; r.x = binop(op, x.x, y.x); r.y = binop(op, x.y, y.y); r.z = binop(op, x.z, y.z); r.w = binop(op, x.w, y.w);
        ((float4 *) d)[g] = r;
    }
}

__device__ __forceinline__ float unop(int op, float x, float p0 = 0.0f, float p1 = 0.0f) {
    switch (op) {
        case B200_SIN:        return sinf(x);
        case B200_COS:        return cosf(x);
        case B200_LOG:        return logf(x);
        case B200_ELU:        return x > 0.0f ? x : expm1f(x);
        case B200_STEP:       return x > 0.0f ? 1.0f : 0.0f;
        case B200_SGN:        return x > 0.0f ? 1.0f : x < 0.0f ? -1.0f : 0.0f;
        case B200_HARDSWISH:  return x * fminf(1.0f, fmaxf(0.0f, (x + 3.0f) / 6.0f));
        case B200_HARDSIGMOID: return fminf(1.0f, fmaxf(0.0f, (x + 3.0f) / 6.0f));
        case B200_LEAKY_RELU: return (x > 0.0f ? x : 0.0f) + p0 * (x < 0.0f ? x : 0.0f);
        case B200_CLAMP:      return x < p0 ? p0 : x > p1 ? p1 : x;
        case B200_SILU:       return x / (1.0f + expf(-x));
        case B200_GELU:       { const float c = 0.044715f, s = 0.79788456080286535587989211986876f; return 0.5f * x * (1.0f + tanhf(s * x * (1.0f + c * x * x))); }
        case B200_RELU:       return fmaxf(x, 0.0f);
        case B200_GELU_QUICK: return x * (1.0f / (1.0f + expf(-1.702f * x)));
        case B200_TANH:       return tanhf(x);
        case B200_SIGMOID:    return 1.0f / (1.0f + expf(-x));
        case B200_GELU_ERF:   return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
        case B200_NEG:        return -x;
        case B200_EXP:        return expf(x);
        case B200_SQR:        return x * x;
        case B200_SQRT:       return sqrtf(x);
        case B200_ABS:        return fabsf(x);
    }
    return x;
}
struct UnArgs { T4 x, dst; int op; float p0, p1; };     // op < 0: scale (y = x*p0 + p1)
__global__ void __launch_bounds__(256) k_unary(const UnArgs A, int64_t total) {
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t) gridDim.x * blockDim.x) {
        int64_t r = g;
        const int64_t i0 = r % A.dst.ne[0]; r /= A.dst.ne[0];
        const int64_t i1 = r % A.dst.ne[1]; r /= A.dst.ne[1];
        const int64_t i2 = r % A.dst.ne[2]; const int64_t i3 = r / A.dst.ne[2];
        const float x = ld_any(A.x.data + i0 * A.x.nb[0] + i1 * A.x.nb[1] + i2 * A.x.nb[2] + i3 * A.x.nb[3], A.x.type);
        const float y = A.op < 0 ? x * A.p0 + A.p1 : unop(A.op, x, A.p0, A.p1);
        st_any(A.dst.data + i0 * A.dst.nb[0] + i1 * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3], A.dst.type, y);
    }
}

struct GluArgs { T4 g, u, dst; int op; int64_t g_off0, u_off0; };   // element offsets along dim 0 (single-tensor form)
__device__ __forceinline__ float gluop(int op, float g) {
    switch (op) {
        case B200_GLU_REGLU:       return fmaxf(g, 0.0f);
        case B200_GLU_GEGLU:       return unop(B200_GELU, g);
        case B200_GLU_SWIGLU:      return g / (1.0f + expf(-g));
        case B200_GLU_GEGLU_ERF:   return unop(B200_GELU_ERF, g);
        case B200_GLU_GEGLU_QUICK: return unop(B200_GELU_QUICK, g);
    }
    return g;
}
__global__ void __launch_bounds__(256) k_glu(const GluArgs A, int64_t total) {
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t) gridDim.x * blockDim.x) {
        int64_t r = g;
        const int64_t i0 = r % A.dst.ne[0]; r /= A.dst.ne[0];
        const int64_t i1 = r % A.dst.ne[1]; r /= A.dst.ne[1];
        const int64_t i2 = r % A.dst.ne[2]; const int64_t i3 = r / A.dst.ne[2];
        const float gv = ld_any(A.g.data + (i0 + A.g_off0) * A.g.nb[0] + i1 * A.g.nb[1] + i2 * A.g.nb[2] + i3 * A.g.nb[3], A.g.type);
        const float uv = ld_any(A.u.data + (i0 + A.u_off0) * A.u.nb[0] + i1 * A.u.nb[1] + i2 * A.u.nb[2] + i3 * A.u.nb[3], A.u.type);
        st_any(A.dst.data + i0 * A.dst.nb[0] + i1 * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3], A.dst.type, gluop(A.op, gv) * uv);
    }
}

// two-tensor F32 form with 16-byte aligned contiguous rows (the ffn of every llama-family graph): 3 x 128-bit accesses per 4 elements
__global__ void __launch_bounds__(256) k_glu_f32x4(const float * __restrict__ g, const float * __restrict__ u, float * __restrict__ d, int op, int64_t n4) {
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t) gridDim.x * blockDim.x) {
        const float4 a = __ldcs((const float4 *) g + i), b = __ldcs((const float4 *) u + i);
        float4 r; r.x = gluop(op, a.x) * b.x; r.y = gluop(op, a.y) * b.y; r.z = gluop(op, a.z) * b.z; r.w = gluop(op, a.w) * b.w;
        ((float4 *) d)[i] = r;
    }
}

// SWIGLU of two contiguous F32 matrices straight into F16 activation tiles (the ffn_down MUL_MAT is the only reader of h): thread = 8 consecutive elements
__global__ void __launch_bounds__(256) k_glu_tiles(const float * __restrict__ g, const float * __restrict__ u, uint8_t * __restrict__ tiles, int op, int64_t n, int64_t k) {
    // thread id -> (row within its group of 8, 8-column chunk, row group): the 8 threads of a row group write one 128-byte core matrix back to back, and 4 consecutive
    // chunks of a row are one 128-byte line of each input (the mapping of k_x_to_f16_tiles)
    const int64_t k8n = k >> 3, total = ((n + 7) >> 3) * 8 * k8n;
    for (int64_t id = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (int64_t) gridDim.x * blockDim.x) {
        const int64_t row = ((id >> 3) / k8n) * 8 + (id & 7), col = ((id >> 3) % k8n) * 8;
        uint4 out = make_uint4(0, 0, 0, 0);
        if (row < n) {
            const float4 a0 = __ldcs((const float4 *) (g + row * k + col)), a1 = __ldcs((const float4 *) (g + row * k + col + 4));
            const float4 b0 = __ldcs((const float4 *) (u + row * k + col)), b1 = __ldcs((const float4 *) (u + row * k + col + 4));
            const __half2 h[4] = { __floats2half2_rn(gluop(op, a0.x) * b0.x, gluop(op, a0.y) * b0.y), __floats2half2_rn(gluop(op, a0.z) * b0.z, gluop(op, a0.w) * b0.w),
                                   __floats2half2_rn(gluop(op, a1.x) * b1.x, gluop(op, a1.y) * b1.y), __floats2half2_rn(gluop(op, a1.z) * b1.z, gluop(op, a1.w) * b1.w) };
            out = *(const uint4 *) h;
        }
        *(uint4 *) (tiles + act_tile_off(row, col, k)) = out;
    }
}

// ================================================================== SOFT_MAX ===================================================
struct SmArgs { T4 x, mask, dst; int has_mask; float scale; int64_t rows; };
// one CTA per row; row kept in shared memory when it fits (<= 12288 floats), else recomputed from global
__global__ void __launch_bounds__(256) k_soft_max(const SmArgs A) {
    extern __shared__ float row[];
    __shared__ float red[32];
    const int64_t r = blockIdx.x;
    const int64_t i1 = r % A.x.ne[1], i2 = (r / A.x.ne[1]) % A.x.ne[2], i3 = r / (A.x.ne[1] * A.x.ne[2]);
    const float * x = (const float *) (A.x.data + i1 * A.x.nb[1] + i2 * A.x.nb[2] + i3 * A.x.nb[3]);
    float * y = (float *) (A.dst.data + i1 * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3]);
    const char * m = A.has_mask ? A.mask.data + i1 * A.mask.nb[1] + (i2 % A.mask.ne[2]) * A.mask.nb[2] + (i3 % A.mask.ne[3]) * A.mask.nb[3] : nullptr;
    const int64_t n = A.x.ne[0];
    float mx = -INFINITY;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        float v = x[i] * A.scale;
        if (m) v += A.mask.type == B200_F16 ? __half2float(((const __half *) m)[i]) : ((const float *) m)[i];
        row[i] = v; mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = (threadIdx.x & 31) < (blockDim.x >> 5) ? red[threadIdx.x & 31] : -INFINITY;
    mx = warp_max(mx);
    __syncthreads();
    float sum = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const float e = expf(row[i] - mx); row[i] = e; sum += e; }
    sum = block_sum(sum, red);
    const float inv = 1.0f / sum;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) y[i] = row[i] * inv;
}

static bool float_type(int t) { return t == B200_F32 || t == B200_F16 || t == B200_BF16; }
// quant_rows.cu
int quant_rows(const b200_tensor * src, const b200_tensor * idx, const b200_tensor * dst, cudaStream_t st);
int dequant_rows(const b200_tensor * src, const b200_tensor * idx, const b200_tensor * dst, cudaStream_t st);

} // namespace b200

using namespace b200;

// ---------------------------------------------------------------------------------------------------------------- C-ABI
static int rms_norm_impl(const b200_tensor * x, const b200_tensor * w, const b200_tensor * add, const b200_tensor * dst, float eps, void * tiles, bool write_f32, void * stream);
extern "C" int b200_rms_norm(const b200_tensor * x, const b200_tensor * w, const b200_tensor * add, const b200_tensor * dst,
                             float eps, void * stream) { return rms_norm_impl(x, w, add, dst, eps, nullptr, true, stream); }
// y = RMS_NORM(x) [* w] handed to a following tensor-core MUL_MAT as prepared F16 tiles in `tiles` (the MUL_MAT's scratch; call it with B200_MM_REUSE_ACT);
// x must be a 2-D [k, n] matrix with k % 8 == 0 and k <= 4096.  dst->data may be NULL: tiles only.
extern "C" int b200_rms_norm_tiles(const b200_tensor * x, const b200_tensor * w, const b200_tensor * dst, void * tiles, float eps, void * stream) {
    if (!x || !dst || !tiles) return B200_ERR_ARG;
    if (x->ne[2] * x->ne[3] != 1 || x->ne[0] % 64 || x->ne[0] > 4096 || x->ne[0] <= 1024 || (uintptr_t) tiles % 16) return B200_ERR_UNSUPPORTED;
    if ((uintptr_t) x->data % 16 || x->nb[1] % 16 || (dst->data && ((uintptr_t) dst->data % 16 || dst->nb[1] % 16))) return B200_ERR_UNSUPPORTED;
    if (w && (w->ne[0] != x->ne[0] || w->ne[1] * w->ne[2] * w->ne[3] != 1 || (uintptr_t) w->data % 16)) return B200_ERR_UNSUPPORTED;
    return rms_norm_impl(x, w, nullptr, dst, eps, tiles, dst->data != nullptr, stream);
}
static int rms_norm_impl(const b200_tensor * x, const b200_tensor * w, const b200_tensor * add, const b200_tensor * dst, float eps, void * tiles, bool write_f32, void * stream) {
    if (!x || !dst) return B200_ERR_ARG;
    if (x->type != B200_F32 || dst->type != B200_F32 || !same_shape(x, dst) || x->nb[0] != 4 || dst->nb[0] != 4) return B200_ERR_UNSUPPORTED;
    if (w   && (w->type   != B200_F32 || w->nb[0]   != 4 || x->ne[0] % w->ne[0]   || x->ne[1] % w->ne[1]   || x->ne[2] % w->ne[2]   || x->ne[3] % w->ne[3]))   return B200_ERR_UNSUPPORTED;
    if (add && (add->type != B200_F32 || add->nb[0] != 4 || x->ne[0] % add->ne[0] || x->ne[1] % add->ne[1] || x->ne[2] % add->ne[2] || x->ne[3] % add->ne[3])) return B200_ERR_UNSUPPORTED;
    const int64_t rows = nrows(x);
    if (rows == 0 || x->ne[0] == 0) return B200_OK;
    NormArgs A; A.x = t4(x); A.dst = t4(dst); A.eps = eps; A.has_w = w != nullptr; A.has_add = add != nullptr; A.rows = rows;
    A.w = w ? t4(w) : A.x; A.add = add ? t4(add) : A.x; A.tiles = (uint8_t *) tiles; A.write_f32 = write_f32 ? 1 : 0;
    if (!write_f32) A.dst = A.x;                                          // (never dereferenced)
    cudaStream_t st = (cudaStream_t) stream;
    if (x->ne[0] <= 1024) k_rms_norm<true><<<(unsigned) ((rows + 7) / 8), 256, 0, st>>>(A);
    else                  k_rms_norm<false><<<(unsigned) rows, 256, 0, st>>>(A);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

static float yarn_corr_dim(int n_dims, int n_ctx_orig, float n_rot, float base) {
    return n_dims * logf(n_ctx_orig / (n_rot * 2 * (float) M_PI)) / (2 * logf(base));
}

extern "C" int b200_rope(const b200_tensor * x, const int32_t * pos, const float * freq_factors, const b200_tensor * dst,
                         const b200_rope_params * p, void * stream) {
    if (!x || !dst || !pos || !p) return B200_ERR_ARG;
    // F16 in / out: the K-shift of the F16 KV cache (ggml_rope_ext_inplace on a cache view, src/llama-kv-cache.cpp build_rope_shift; CPU ggml_compute_forward_rope_f16)
    if ((x->type != B200_F32 && x->type != B200_F16) || dst->type != x->type || !same_shape(x, dst) || x->nb[0] != type_size(x->type) || dst->nb[0] != x->nb[0]) return B200_ERR_UNSUPPORTED;
    if ((p->mode != 0 && p->mode != 2) || p->n_dims <= 0 || p->n_dims % 2 || p->n_dims > x->ne[0] || x->ne[0] % 2) return B200_ERR_UNSUPPORTED;
    RopeArgs A; A.x = t4(x); A.dst = t4(dst); A.pos = pos; A.ff = freq_factors; A.n_dims = p->n_dims; A.mode = p->mode;
    A.theta_scale = powf(p->freq_base, -2.0f / p->n_dims);
    A.freq_scale = p->freq_scale; A.ext_factor = p->ext_factor; A.attn_factor = p->attn_factor;
    const float lo = floorf(yarn_corr_dim(p->n_dims, p->n_ctx_orig, p->beta_fast, p->freq_base));
    const float hi = ceilf (yarn_corr_dim(p->n_dims, p->n_ctx_orig, p->beta_slow, p->freq_base));
    A.corr0 = lo < 0 ? 0 : lo; A.corr1 = hi > p->n_dims - 1 ? p->n_dims - 1 : hi;
    const int64_t total = nelem(x) / 2;
    if (total == 0) return B200_OK;
    const bool rows = p->n_dims / 2 <= ROPE_TAB && x->ne[2] <= 0x7fffffff && x->ne[3] <= 65535 && x->ne[1] * x->ne[0] >= 512;
    const dim3 rgrid((unsigned) x->ne[2], (unsigned) x->ne[3]);
    if (x->type == B200_F32) {
        if (rows) k_rope_rows<float><<<rgrid, 256, 0, (cudaStream_t) stream>>>(A);
        else      k_rope<float><<<grid_for(total, 256), 256, 0, (cudaStream_t) stream>>>(A, total);
    } else {
        if (rows) k_rope_rows<__half><<<rgrid, 256, 0, (cudaStream_t) stream>>>(A);
        else      k_rope<__half><<<grid_for(total, 256), 256, 0, (cudaStream_t) stream>>>(A, total);
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_set_rows(const b200_tensor * src, const b200_tensor * idx, const b200_tensor * dst, void * stream) {
    if (!src || !idx || !dst) return B200_ERR_ARG;
    const bool qdst = dst->type == B200_Q8_0 || dst->type == B200_Q4_0;      // quantised KV cache (-ctk / -ctv): quant_rows.cu
    if (src->type != B200_F32 || (!float_type(dst->type) && !qdst) || (idx->type != B200_I64 && idx->type != B200_I32)) return B200_ERR_UNSUPPORTED;
    if (src->ne[0] != dst->ne[0] || src->ne[2] != dst->ne[2] || src->ne[3] != dst->ne[3] || idx->ne[0] != src->ne[1]) return B200_ERR_UNSUPPORTED;
    if (idx->ne[1] == 0 || idx->ne[2] == 0 || src->ne[2] % idx->ne[1] || src->ne[3] % idx->ne[2]) return B200_ERR_UNSUPPORTED;
    if (qdst) return quant_rows(src, idx, dst, (cudaStream_t) stream);
    const int64_t total = nelem(src);
    if (total == 0) return B200_OK;
    RowsArgs A = { t4(src), t4(idx), t4(dst) };
    k_set_rows<<<grid_for(total, 256), 256, 0, (cudaStream_t) stream>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_get_rows(const b200_tensor * src, const b200_tensor * idx, const b200_tensor * dst, void * stream) {
    if (!src || !idx || !dst) return B200_ERR_ARG;
    if (idx->type != B200_I32 || src->ne[0] != dst->ne[0] || dst->ne[1] != idx->ne[0] || dst->ne[2] != idx->ne[1] || dst->ne[3] != idx->ne[2]) return B200_ERR_UNSUPPORTED;
    if (is_quant(src->type)) return dst->type == B200_F32 ? dequant_rows(src, idx, dst, (cudaStream_t) stream) : B200_ERR_UNSUPPORTED;      // quantised token_embd
    if (!float_type(src->type) || !float_type(dst->type)) return B200_ERR_UNSUPPORTED;
    const int64_t total = nelem(dst);
    if (total == 0) return B200_OK;
    RowsArgs A = { t4(src), t4(idx), t4(dst) };
    k_get_rows<<<grid_for(total, 256), 256, 0, (cudaStream_t) stream>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_cpy(const b200_tensor * src, const b200_tensor * dst, void * stream) {
    if (!src || !dst) return B200_ERR_ARG;
    const bool ints = src->type == B200_I32 && dst->type == B200_I32;
    // F32 -> q8_0 / q4_0 and quant -> F32 / F16 between tensors of the same shape (the K-shift of a quantised KV cache casts through F32)
    if (src->type == B200_F32 && (dst->type == B200_Q8_0 || dst->type == B200_Q4_0)) return same_shape(src, dst) ? quant_rows(src, nullptr, dst, (cudaStream_t) stream) : B200_ERR_UNSUPPORTED;
    if (is_quant(src->type) && (dst->type == B200_F32 || dst->type == B200_F16))     return same_shape(src, dst) ? dequant_rows(src, nullptr, dst, (cudaStream_t) stream) : B200_ERR_UNSUPPORTED;
    if (!ints && (!float_type(src->type) || !float_type(dst->type))) return B200_ERR_UNSUPPORTED;
    const int64_t total = nelem(src);
    if (total != nelem(dst)) return B200_ERR_UNSUPPORTED;
    if (total == 0) return B200_OK;
    cudaStream_t st = (cudaStream_t) stream;
    if (is_contig(src) && is_contig(dst) && ((uintptr_t) src->data % 16 == 0) && ((uintptr_t) dst->data % 16 == 0)) {
        if (src->type == dst->type) { B200_CUDA_TRY(cudaMemcpyAsync(dst->data, src->data, (size_t) total * type_size(src->type), cudaMemcpyDeviceToDevice, st)); return B200_OK; }
        const unsigned g = grid_for(total / 4, 256);
        if (src->type == B200_F32 && dst->type == B200_F16) { k_cpy_contig<float, __half><<<g, 256, 0, st>>>((const float *) src->data, (__half *) dst->data, total); B200_LAUNCH_CHECK(); return B200_OK; }
        if (src->type == B200_F16 && dst->type == B200_F32) { k_cpy_contig<__half, float><<<g, 256, 0, st>>>((const __half *) src->data, (float *) dst->data, total); B200_LAUNCH_CHECK(); return B200_OK; }
    }
    CpyArgs A = { t4(src), t4(dst) };
    k_cpy<<<grid_for(total, 256), 256, 0, st>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_binary(int op, const b200_tensor * a, const b200_tensor * b, const b200_tensor * dst, void * stream) {
    if (!a || !b || !dst) return B200_ERR_ARG;
    if (op < B200_ADD || op > B200_DIV || !float_type(a->type) || !float_type(b->type) || !float_type(dst->type) || !same_shape(a, dst)) return B200_ERR_UNSUPPORTED;
    for (int i = 0; i < 4; ++i) if (b->ne[i] == 0 || a->ne[i] % b->ne[i]) return nelem(dst) == 0 ? B200_OK : B200_ERR_UNSUPPORTED;
    const int64_t total = nelem(dst);
    if (total == 0) return B200_OK;
    cudaStream_t st = (cudaStream_t) stream;
    const bool all32 = a->type == B200_F32 && b->type == B200_F32 && dst->type == B200_F32;
    const int64_t nb_ = nelem(b);
    const bool b_rowlike = is_contig(b) && (same_shape(a, b) || (b->ne[0] == a->ne[0] && nb_ == b->ne[0]));
    if (all32 && is_contig(a) && is_contig(dst) && b_rowlike && a->ne[0] % 4 == 0 &&
        (((uintptr_t) a->data | (uintptr_t) b->data | (uintptr_t) dst->data) % 16 == 0)) {
        k_binary_f32x4<<<grid_for(total / 4, 256), 256, 0, st>>>((const float *) a->data, (const float *) b->data, (float *) dst->data, total / 4, nb_ / 4, op);
    } else {
        BinArgs A = { t4(a), t4(b), t4(dst), op };
        k_binary<<<grid_for(total, 256), 256, 0, st>>>(A, total);
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_unary(int op, const b200_tensor * x, const b200_tensor * dst, void * stream) {
    if (!x || !dst) return B200_ERR_ARG;
    if (op < B200_SILU || op > B200_ABS || !float_type(x->type) || !float_type(dst->type) || !same_shape(x, dst)) return B200_ERR_UNSUPPORTED;
    const int64_t total = nelem(dst);
    if (total == 0) return B200_OK;
    UnArgs A = { t4(x), t4(dst), op, 0.0f, 0.0f };
    k_unary<<<grid_for(total, 256), 256, 0, (cudaStream_t) stream>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_unary_param(int op, const b200_tensor * x, const b200_tensor * dst, float p0, float p1, void * stream) {
    if (!x || !dst) return B200_ERR_ARG;
    if (op < B200_SILU || op > B200_CLAMP || !float_type(x->type) || !float_type(dst->type) || !same_shape(x, dst)) return B200_ERR_UNSUPPORTED;
    const int64_t total = nelem(dst);
    if (total == 0) return B200_OK;
    UnArgs A = { t4(x), t4(dst), op, p0, p1 };
    k_unary<<<grid_for(total, 256), 256, 0, (cudaStream_t) stream>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_scale(const b200_tensor * x, const b200_tensor * dst, float scale, float bias, void * stream) {
    if (!x || !dst) return B200_ERR_ARG;
    if (!float_type(x->type) || !float_type(dst->type) || !same_shape(x, dst)) return B200_ERR_UNSUPPORTED;
    const int64_t total = nelem(dst);
    if (total == 0) return B200_OK;
    UnArgs A = { t4(x), t4(dst), -1, scale, bias };
    k_unary<<<grid_for(total, 256), 256, 0, (cudaStream_t) stream>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_glu(int op, const b200_tensor * gate_or_x, const b200_tensor * up, const b200_tensor * dst, int swapped, void * stream) {
    if (!gate_or_x || !dst) return B200_ERR_ARG;
    if (!(op == B200_GLU_REGLU || op == B200_GLU_GEGLU || op == B200_GLU_SWIGLU || op == B200_GLU_GEGLU_ERF || op == B200_GLU_GEGLU_QUICK)) return B200_ERR_UNSUPPORTED;
    if (!float_type(gate_or_x->type) || !float_type(dst->type)) return B200_ERR_UNSUPPORTED;
    GluArgs A; A.op = op; A.dst = t4(dst); A.g = t4(gate_or_x); A.g_off0 = 0; A.u_off0 = 0;
    if (up) {
        if (!float_type(up->type) || !same_shape(gate_or_x, up) || !same_shape(up, dst)) return B200_ERR_UNSUPPORTED;
        A.u = t4(up);
    } else {
        if (gate_or_x->ne[0] != 2 * dst->ne[0] || gate_or_x->ne[1] != dst->ne[1] || gate_or_x->ne[2] != dst->ne[2] || gate_or_x->ne[3] != dst->ne[3]) return B200_ERR_UNSUPPORTED;
        A.u = A.g;
        if (swapped) A.g_off0 = dst->ne[0]; else A.u_off0 = dst->ne[0];
    }
    const int64_t total = nelem(dst);
    if (total == 0) return B200_OK;
    if (up && gate_or_x->type == B200_F32 && up->type == B200_F32 && dst->type == B200_F32 && is_contig(gate_or_x) && is_contig(up) && is_contig(dst) &&
        total % 4 == 0 && (((uintptr_t) gate_or_x->data | (uintptr_t) up->data | (uintptr_t) dst->data) % 16) == 0) {
        k_glu_f32x4<<<grid_for(total / 4, 256), 256, 0, (cudaStream_t) stream>>>((const float *) gate_or_x->data, (const float *) up->data, (float *) dst->data, op, total / 4);
        B200_LAUNCH_CHECK();
        return B200_OK;
    }
    k_glu<<<grid_for(total, 256), 256, 0, (cudaStream_t) stream>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// h = GLU(gate, up) of two contiguous [k, n] F32 matrices as prepared F16 tiles for the MUL_MAT that follows (ffn_down); see b200_rms_norm_tiles
extern "C" int b200_glu_tiles(int op, const b200_tensor * gate, const b200_tensor * up, void * tiles, void * stream) {
    if (!gate || !up || !tiles) return B200_ERR_ARG;
    if (gate->type != B200_F32 || up->type != B200_F32 || !same_shape(gate, up) || !is_contig(gate) || !is_contig(up) || gate->ne[2] * gate->ne[3] != 1) return B200_ERR_UNSUPPORTED;
    if (gate->ne[0] % 64 || (((uintptr_t) gate->data | (uintptr_t) up->data | (uintptr_t) tiles) % 16)) return B200_ERR_UNSUPPORTED;
    if (op != B200_GLU_REGLU && op != B200_GLU_GEGLU && op != B200_GLU_SWIGLU && op != B200_GLU_GEGLU_ERF && op != B200_GLU_GEGLU_QUICK) return B200_ERR_UNSUPPORTED;
    const int64_t n = gate->ne[1], k = gate->ne[0];
    if (n == 0) return B200_OK;
    k_glu_tiles<<<grid_for(((n + 7) >> 3) * 8 * (k >> 3), 256), 256, 0, (cudaStream_t) stream>>>((const float *) gate->data, (const float *) up->data, (uint8_t *) tiles, op, n, k);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_soft_max(const b200_tensor * x, const b200_tensor * mask, const b200_tensor * dst, float scale, float max_bias, void * stream) {
    if (!x || !dst) return B200_ERR_ARG;
    if (x->type != B200_F32 || dst->type != B200_F32 || !same_shape(x, dst) || x->nb[0] != 4 || dst->nb[0] != 4 || max_bias != 0.0f) return B200_ERR_UNSUPPORTED;
    if (mask && ((mask->type != B200_F32 && mask->type != B200_F16) || mask->ne[0] != x->ne[0] || mask->ne[1] < x->ne[1] ||
                 mask->ne[2] == 0 || mask->ne[3] == 0 || x->ne[2] % mask->ne[2] || x->ne[3] % mask->ne[3])) return B200_ERR_UNSUPPORTED;
    if (x->ne[0] > 24576) return B200_ERR_UNSUPPORTED;
    static smem_mask_t sm_attr{0};
    B200_CUDA_TRY(ensure_dyn_smem(k_soft_max, 24576 * 4, sm_attr));
    const int64_t rows = nrows(x);
    if (rows == 0 || x->ne[0] == 0) return B200_OK;
    SmArgs A; A.x = t4(x); A.dst = t4(dst); A.has_mask = mask != nullptr; A.mask = mask ? t4(mask) : A.x; A.scale = scale; A.rows = rows;
    k_soft_max<<<(unsigned) rows, 256, (size_t) x->ne[0] * 4, (cudaStream_t) stream>>>(A);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// ================================================================== ops of the APM (Whisper) / VPM (SigLip) encoder graphs ======================
// without them every encoder layer bounced to the CPU backend through the scheduler.
//   NORM    ggml_norm       tools/omni/audition.cpp:449,653,674, vision.cpp:546   (CPU: ggml-cpu/ops.cpp:3450-3495, CUDA: norm.cu:5)
//   IM2COL  ggml_conv_1d/2d tools/omni/audition.cpp:379,384, vision.cpp:519       (CPU: ops.cpp:6160-6301, CUDA: im2col.cu:6)
//   POOL_1D ggml_pool_1d    tools/omni/audition.cpp:697                            (CPU: ops.cpp:7212-7260; the reference CUDA backend has no POOL_1D)
namespace b200 {

// LayerNorm without affine part: y = (x - mean) / sqrt(var + eps), var = mean((x - mean)^2) — two passes over the row like the CPU (mean first, then
// the centred sum of squares), the row staged in registers when it fits
template <bool WARP_ROWS>
__global__ void __launch_bounds__(256) k_norm(const NormArgs A) {
    __shared__ float red[32];
    const int lane = threadIdx.x & 31;
    const int64_t row = WARP_ROWS ? (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) : blockIdx.x;
    if (WARP_ROWS && row >= A.rows) return;
    const int64_t i1 = row % A.x.ne[1], i2 = (row / A.x.ne[1]) % A.x.ne[2], i3 = row / (A.x.ne[1] * A.x.ne[2]);
    const float * x = (const float *) (A.x.data + i1 * A.x.nb[1] + i2 * A.x.nb[2] + i3 * A.x.nb[3]);
    float * y = (float *) (A.dst.data + i1 * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3]);
    const int64_t n = A.x.ne[0];
    const int tid = WARP_ROWS ? lane : threadIdx.x, nt = WARP_ROWS ? 32 : blockDim.x;
    float s = 0.0f;
    for (int64_t i = tid; i < n; i += nt) s += x[i];
    s = WARP_ROWS ? warp_sum(s) : block_sum(s, red);
    const float mean = s / (float) n;
    float ss = 0.0f;
    for (int64_t i = tid; i < n; i += nt) { const float d = x[i] - mean; ss += d * d; }
    ss = WARP_ROWS ? warp_sum(ss) : block_sum(ss, red);
    const float scale = 1.0f / sqrtf(ss / (float) n + A.eps);
    for (int64_t i = tid; i < n; i += nt) y[i] = __fmul_rn(x[i] - mean, scale);
}

struct Im2colArgs {
    const char * x; char * dst; int dst_f16;
    int64_t N, IC, IH, IW, KH, KW, OH, OW; int64_t x_nb_n, x_nb_c, x_nb_h;     // byte strides of the F32 input: batch, channel, row
    int s0, s1, p0, p1, d0, d1;
};
// thread = one element of dst [N][OH][OW][IC*KH*KW] (the innermost index is the contiguous one: coalesced stores; the gathers hit a KW-wide window)
__global__ void __launch_bounds__(256) k_im2col(const Im2colArgs A, int64_t total) {
    const int64_t ckk = A.IC * A.KH * A.KW;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t) gridDim.x * blockDim.x) {      // grid_for caps the grid
    const int64_t c = i % ckk, r = i / ckk;
    const int64_t ikw = c % A.KW, ikh = (c / A.KW) % A.KH, iic = c / (A.KW * A.KH);
    const int64_t iow = r % A.OW, ioh = (r / A.OW) % A.OH, in = r / (A.OW * A.OH);
    const int64_t iiw = iow * A.s0 + ikw * A.d0 - A.p0, iih = ioh * A.s1 + ikh * A.d1 - A.p1;
    float v = 0.0f;
    if (iih >= 0 && iih < A.IH && iiw >= 0 && iiw < A.IW) v = *(const float *) (A.x + in * A.x_nb_n + iic * A.x_nb_c + iih * A.x_nb_h + iiw * 4);
    if (A.dst_f16) ((__half *) A.dst)[i] = __float2half_rn(v); else ((float *) A.dst)[i] = v;
    }
}

struct PoolArgs { const char * x; float * dst; int x_f16, op, k; int64_t rs, x_row_bytes; };
// thread = one output element: rows of the source are nb[1] apart and walked back to back, exactly as the CPU loop does (k == stride, no padding)
__global__ void __launch_bounds__(256) k_pool_1d(const PoolArgs A, int64_t total) {
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t) gridDim.x * blockDim.x) {
    const int64_t row = i / A.rs, o = i % A.rs;
    const char * src = A.x + row * A.x_row_bytes;
    float acc = A.op == 0 ? -3.402823466e+38f : 0.0f;                     // GGML_OP_POOL_MAX = 0, GGML_OP_POOL_AVG = 1
    for (int ki = 0; ki < A.k; ++ki) {
        const int64_t j = o * A.k + ki;
        const float v = A.x_f16 ? __half2float(((const __half *) src)[j]) : ((const float *) src)[j];
        if (A.op == 0) { if (v > acc) acc = v; } else acc += v;
    }
    A.dst[i] = A.op == 0 ? acc : __fdiv_rn(acc, (float) A.k);
    }
}

} // namespace b200

extern "C" int b200_norm(const b200_tensor * x, const b200_tensor * dst, float eps, void * stream) {
    if (!x || !dst) return B200_ERR_ARG;
    if (x->type != B200_F32 || dst->type != B200_F32 || !same_shape(x, dst) || x->nb[0] != 4 || dst->nb[0] != 4) return B200_ERR_UNSUPPORTED;
    const int64_t rows = nrows(x);
    if (rows == 0 || x->ne[0] == 0) return B200_OK;
    NormArgs A; A.x = t4(x); A.dst = t4(dst); A.eps = eps; A.has_w = 0; A.has_add = 0; A.rows = rows; A.w = A.x; A.add = A.x; A.tiles = nullptr; A.write_f32 = 1;
    cudaStream_t st = (cudaStream_t) stream;
    if (x->ne[0] <= 1024) k_norm<true><<<(unsigned) ((rows + 7) / 8), 256, 0, st>>>(A);
    else                  k_norm<false><<<(unsigned) rows, 256, 0, st>>>(A);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_im2col(const b200_tensor * kernel, const b200_tensor * x, const b200_tensor * dst, int s0, int s1, int p0, int p1, int d0, int d1,
                           int is_2d, void * stream) {
    if (!kernel || !x || !dst) return B200_ERR_ARG;
    if (x->type != B200_F32 || x->nb[0] != 4 || (dst->type != B200_F16 && dst->type != B200_F32) || !is_contig(dst)) return B200_ERR_UNSUPPORTED;
    Im2colArgs A;
    A.x = (const char *) x->data; A.dst = (char *) dst->data; A.dst_f16 = dst->type == B200_F16;
    A.N = is_2d ? x->ne[3] : x->ne[2]; A.IC = is_2d ? x->ne[2] : x->ne[1]; A.IH = is_2d ? x->ne[1] : 1; A.IW = x->ne[0];
    A.KH = is_2d ? kernel->ne[1] : 1; A.KW = kernel->ne[0]; A.OH = is_2d ? dst->ne[2] : 1; A.OW = dst->ne[1];
    A.x_nb_n = is_2d ? x->nb[3] : x->nb[2]; A.x_nb_c = is_2d ? x->nb[2] : x->nb[1]; A.x_nb_h = is_2d ? x->nb[1] : 0;
    A.s0 = s0; A.s1 = s1; A.p0 = p0; A.p1 = p1; A.d0 = d0; A.d1 = d1;
    if (dst->ne[0] != A.IC * A.KH * A.KW) return B200_ERR_ARG;
    const int64_t total = nelem(dst);
    if (total == 0) return B200_OK;
    if (total / 256 + 1 > 0x7fffffffll) return B200_ERR_UNSUPPORTED;
    k_im2col<<<grid_for(total, 256), 256, 0, (cudaStream_t) stream>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_pool_1d(const b200_tensor * x, const b200_tensor * dst, int op, int k0, int s0, int p0, void * stream) {
    if (!x || !dst) return B200_ERR_ARG;
    if ((x->type != B200_F32 && x->type != B200_F16) || dst->type != B200_F32 || !is_contig(dst) || !is_contig(x)) return B200_ERR_UNSUPPORTED;
    if (k0 != s0 || p0 != 0 || k0 <= 0 || (op != 0 && op != 1)) return B200_ERR_UNSUPPORTED;      // what the reference implements too (ops.cpp:7264-7280)
    const int64_t total = nelem(dst);
    if (total == 0) return B200_OK;
    PoolArgs A; A.x = (const char *) x->data; A.dst = (float *) dst->data; A.x_f16 = x->type == B200_F16; A.op = op; A.k = k0; A.rs = dst->ne[0]; A.x_row_bytes = x->nb[1];
    k_pool_1d<<<grid_for(total, 256), 256, 0, (cudaStream_t) stream>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
