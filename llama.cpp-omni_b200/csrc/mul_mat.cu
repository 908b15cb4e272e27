// mul_mat.cu — GGML_OP_MUL_MAT entry point (replaces ggml_cuda_mul_mat dispatcher, ggml-cuda.cu:2001-2084).
//   quantised W : activations -> q8 records (quant_act.cu) -> mmvq (<= 8 columns per pass)
//   f32/f16/bf16 W: mmvf-style warp-per-row kernel below (replaces mul_mat_vec_f, ggml-cuda/mmvf.cu); activations are rounded to
//                   the weight type first exactly like the CPU oracle does (vec_dot_type, ggml-cpu.c:196-350)
//                   more than 8 columns off the tensor-core path (k not a multiple of 64, F32 x F32 attention products of the omni encoders, F16 activations
//                   of ggml_conv_1d / conv_2d): k_mm_simt, a shared-memory-tiled F32 GEMM, ONE launch over all batch slices
// Batch dims follow ggml broadcast rules: w.ne[2] divides x.ne[2], w.ne[3] divides x.ne[3].
#include "common.cuh"
#include <type_traits>
#include <stdlib.h>

namespace b200 {

int matvec_q_cols(const void * w, int type, int layout, int64_t m, int64_t row_stride, const void * act, int64_t k, int ncols,
                  float * y, int64_t y_col_stride, cudaStream_t st);
// mmq_tc.cu: tcgen05 dequant-GEMM for n > 8 columns
bool   mmq_tc_supported(int type, int layout, int64_t k, int64_t n, const void * w, int64_t row_stride);
size_t mmq_tc_scratch_bytes(int64_t k, int64_t n);
int    mmq_tc(const void * w, int type, int64_t m, int64_t k, const float * x, int64_t x_ld, int64_t n, float * dst, int64_t dst_ld, void * scratch, bool reuse_tiles, cudaStream_t st, const float * resid, bool * resid_fused);
bool   mmq_tc_fuses_resid(int64_t m, int64_t k, int64_t n);
bool   mmq_tc_multi_ok(int nseg, const int * type, const int64_t * m, int64_t k, int64_t n);
int    mmq_tc_multi(int nseg, const void * const * w, const int * type, const int64_t * m, int64_t k, const float * x, int64_t x_ld, int64_t n, float * const * dst, const int64_t * dst_ld,
                    void * scratch, bool reuse_tiles, cudaStream_t st, const float * resid, bool * resid_fused);
bool   mm_f16_tc_supported(int type, int64_t k, int64_t n, const void * w, int64_t row_stride);
int    mm_f16_tc(const void * w, int64_t m, int64_t k, const float * x, int64_t x_ld, int64_t n, float * dst, int64_t dst_ld, void * scratch, bool reuse_tiles, cudaStream_t st);
static bool tc_disabled() { static const bool off = getenv("B200_DISABLE_TC") && atoi(getenv("B200_DISABLE_TC")) != 0; return off; }
// B200_NO_SIMT_GEMM=1: float weights with more than 8 columns off the tensor-core path go back to one k_mmvf launch per 8 columns (A/B switch for k_mm_simt; read per call)
static bool simt_disabled() { const char * e = getenv("B200_NO_SIMT_GEMM"); return e && atoi(e) != 0; }

template <typename WT> __device__ __forceinline__ float w2f(WT v);
template <> __device__ __forceinline__ float w2f<float>(float v) { return v; }
template <> __device__ __forceinline__ float w2f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float w2f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename WT> __device__ __forceinline__ float round_like(float v);
template <> __device__ __forceinline__ float round_like<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_like<__half>(float v) { return __half2float(__float2half_rn(v)); }
template <> __device__ __forceinline__ float round_like<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

struct MmvfArgs {
    const char * w; const char * x; char * y;
    int64_t m, k, n;
    int64_t w_nb1, w_nb2, w_nb3, x_nb1, x_nb2, x_nb3, y_nb1, y_nb2, y_nb3;
    int64_t ne2, ne3, r2, r3;        // dst batch dims and broadcast ratios
};

// one warp per (row, batch); up to NC columns per pass; 16-byte loads when the row is aligned, scalar otherwise
template <typename WT, int NC>
__global__ void __launch_bounds__(256) k_mmvf(const MmvfArgs A, int64_t col0) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t rows_total = A.m * A.ne2 * A.ne3;
    if (warp >= rows_total) return;
    const int64_t r = warp % A.m, i2 = (warp / A.m) % A.ne2, i3 = warp / (A.m * A.ne2);
    const WT * w = (const WT *) (A.w + r * A.w_nb1 + (i2 / A.r2) * A.w_nb2 + (i3 / A.r3) * A.w_nb3);
    const char * xb = A.x + i2 * A.x_nb2 + i3 * A.x_nb3;
    float acc[NC];
#pragma unroll
    for (int n = 0; n < NC; ++n) acc[n] = 0.0f;
    constexpr int VEC = 16 / sizeof(WT);
    const bool vec_ok = ((uintptr_t) w % 16 == 0) && (A.k % VEC == 0) && ((uintptr_t) xb % 16 == 0) && (A.x_nb1 % 16 == 0);
    if (vec_ok) {
        for (int64_t i = (int64_t) lane * VEC; i < A.k; i += 32 * VEC) {
            const uint4 raw = ldg_stream16(w + i);
            const WT * wv = (const WT *) &raw;
#pragma unroll
            for (int n = 0; n < NC; ++n) {
                if (col0 + n < A.n) {
                    const float * x = (const float *) (xb + (col0 + n) * A.x_nb1) + i;
#pragma unroll
                    for (int e = 0; e < VEC; e += 4) {
                        const float4 xv = *(const float4 *) (x + e);
                        acc[n] += w2f<WT>(wv[e]) * round_like<WT>(xv.x) + w2f<WT>(wv[e + 1]) * round_like<WT>(xv.y)
                                + w2f<WT>(wv[e + 2]) * round_like<WT>(xv.z) + w2f<WT>(wv[e + 3]) * round_like<WT>(xv.w);
                    }
                }
            }
        }
    } else {
        for (int64_t i = lane; i < A.k; i += 32) {
            const float wv = w2f<WT>(w[i]);
#pragma unroll
            for (int n = 0; n < NC; ++n)
                if (col0 + n < A.n) acc[n] += wv * round_like<WT>(((const float *) (xb + (col0 + n) * A.x_nb1))[i]);
        }
    }
#pragma unroll
    for (int n = 0; n < NC; ++n) {
        const float v = warp_sum(acc[n]);
        if (lane == 0 && col0 + n < A.n) *(float *) (A.y + r * 4 + (col0 + n) * A.y_nb1 + i2 * A.y_nb2 + i3 * A.y_nb3) = v;
    }
}

// F16 weights, ONE activation column, 2-D weight: the decode matvec of an F16 model (BASELINE.json configs[4], the TTS llama, the projector) — HBM-bound.
// The CTA rounds the activation vector to F16 once into shared memory (the oracle's vec_dot_type for F16 weights: ggml-cpu.c type_traits_cpu[F16], products of F16
// values accumulated in F32); every warp then streams TWO rows at a time with two 16-byte loads per row in flight (4 x 512 B per warp on the wire), so that
// ~128 KB per SM are outstanding — what 6.4 TB/s x ~1 us of latency needs.  Replaces mul_mat_vec_f (ggml-cuda/mmvf.cu:309) for ncols = 1.
__device__ __forceinline__ float dot8_f16(const uint4 w, const uint4 x, float acc) {
    const __half2 * wh = (const __half2 *) &w, * xh = (const __half2 *) &x;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 a = __half22float2(wh[i]), b = __half22float2(xh[i]); acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); }
    return acc;
}
// GLU: the two streamed rows are row r of `w` (gate) and row r of `w2` (up), y[r] = silu(gate . x) * (up . x) — the arithmetic of the separate GLU kernel (ops_misc.cu gluop);
// otherwise rows r0, r0 + 1 of `w`, y = W . x (+ resid: the ADD that follows wo / ffn_down; resid may be y)
template <bool GLU>
__global__ void __launch_bounds__(256) k_mmvf16_stream(const __half * __restrict__ w, const __half * __restrict__ w2, int64_t w_ld, const float * __restrict__ x, float * y,
                                                       const float * resid, int64_t m, int64_t k) {
    extern __shared__ __align__(16) __half xs[];
    for (int64_t i = (int64_t) threadIdx.x * 4; i < k; i += 256 * 4) {
        const float4 v = *(const float4 *) (x + i);
        *(__half2 *) (xs + i) = __floats2half2_rn(v.x, v.y); *(__half2 *) (xs + i + 2) = __floats2half2_rn(v.z, v.w);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t) gridDim.x * 8;
    constexpr int RPW = GLU ? 1 : 2;                                       // output rows per warp pass
    for (int64_t r0 = ((int64_t) blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW; r0 < m; r0 += nwarps * RPW) {
        const bool two = !GLU && r0 + 1 < m;
        const __half * w0 = w + r0 * w_ld, * w1 = GLU ? w2 + r0 * w_ld : w + (two ? r0 + 1 : r0) * w_ld;
        float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
        int64_t i = (int64_t) lane * 8;
        for (; i + 256 < k; i += 512) {
            const uint4 p0 = ldg_stream16(w0 + i), p1 = ldg_stream16(w1 + i), q0 = ldg_stream16(w0 + i + 256), q1 = ldg_stream16(w1 + i + 256);
            const uint4 xa = *(const uint4 *) (xs + i), xb = *(const uint4 *) (xs + i + 256);
            a0 = dot8_f16(p0, xa, a0); a1 = dot8_f16(p1, xa, a1); b0 = dot8_f16(q0, xb, b0); b1 = dot8_f16(q1, xb, b1);
        }
        if (i < k) {
            const uint4 p0 = ldg_stream16(w0 + i), p1 = ldg_stream16(w1 + i);
            const uint4 xa = *(const uint4 *) (xs + i);
            a0 = dot8_f16(p0, xa, a0); a1 = dot8_f16(p1, xa, a1);
        }
        const float s0 = warp_sum(a0 + b0), s1 = warp_sum(a1 + b1);
        if (lane == 0) {
            if (GLU) y[r0] = (s0 / (1.0f + expf(-s0))) * s1;
            else if (resid) { const float q0 = resid[r0], q1 = two ? resid[r0 + 1] : 0.0f; y[r0] = s0 + q0; if (two) y[r0 + 1] = s1 + q1; }
            else { y[r0] = s0; if (two) y[r0 + 1] = s1; }
        }
    }
}

// shapes the streaming F16 matvec takes: one column, 2-D weight with 16-byte aligned rows, k a multiple of 256 that fits shared memory as F16
static bool mmvf16_stream_ok(const MmvfArgs & A) {
    return A.n == 1 && A.ne2 * A.ne3 == 1 && A.k % 256 == 0 && A.k <= 24576 && A.w_nb1 % 16 == 0 && ((uintptr_t) A.w | (uintptr_t) A.x) % 16 == 0 && A.m >= 64;
}
static int run_mmvf16_stream(const MmvfArgs & A, const void * w_up, const float * resid, cudaStream_t st) {
    const int64_t cap = (int64_t) sm_count() * 8;
    if (w_up) {
        int64_t g = (A.m + 7) / 8; if (g > cap) g = cap;
        k_mmvf16_stream<true><<<(unsigned) g, 256, (size_t) A.k * 2, st>>>((const __half *) A.w, (const __half *) w_up, A.w_nb1 / 2, (const float *) A.x, (float *) A.y, nullptr, A.m, A.k);
    } else {
        int64_t g = (A.m / 2 + 7) / 8; if (g > cap) g = cap;
        k_mmvf16_stream<false><<<(unsigned) g, 256, (size_t) A.k * 2, st>>>((const __half *) A.w, nullptr, A.w_nb1 / 2, (const float *) A.x, (float *) A.y, resid, A.m, A.k);
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

template <typename WT>
static int run_mmvf(const MmvfArgs & A, cudaStream_t st) {
    if (std::is_same<WT, __half>::value && mmvf16_stream_ok(A)) return run_mmvf16_stream(A, nullptr, nullptr, st);
    const int64_t warps = A.m * A.ne2 * A.ne3;
    const unsigned grid = (unsigned) ((warps + 7) / 8);
    for (int64_t c0 = 0; c0 < A.n; c0 += 8) {
        const int64_t nc = A.n - c0;
        if      (nc == 1) k_mmvf<WT, 1><<<grid, 256, 0, st>>>(A, c0);
        else if (nc == 2) k_mmvf<WT, 2><<<grid, 256, 0, st>>>(A, c0);
        else if (nc <= 4) k_mmvf<WT, 4><<<grid, 256, 0, st>>>(A, c0);
        else              k_mmvf<WT, 8><<<grid, 256, 0, st>>>(A, c0);
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// dst[m, n] = sum_k W[m, k] . X[n, k] for float weights and MORE than 8 columns when the tcgen05 GEMM does not apply (k not a multiple of 64, unaligned rows, F32 or BF16
// weights) — the attention products of the omni encoders are exactly that: SigLip's K.Q^T / V.softmax are F32 x F32 with k = 72 or n_pos over 16 heads
// (tools/omni/vision.cpp:662-668), Whisper's V.softmax has k = 50 . (iter + 1) (tools/omni/audition.cpp:620-624), and ggml_conv_1d / conv_2d multiply an F16 im2col matrix
// by an F16 kernel (ggml.c: ggml_conv_1d -> ggml_mul_mat(im2col, kernel)), i.e. F16 ACTIVATIONS (XT = __half).  Before this kernel those ran as one k_mmvf launch per 8
// columns (2048 launches per attention product of a 1024-patch frame).  64 x 64 output tile per CTA, K in steps of 32 through shared memory, 4 x 4 outputs per thread,
// F32 FMA: the CPU oracle's arithmetic (activations rounded to the weight type = vec_dot_type, F32 accumulation), only the summation order differs.  All batch slices in ONE
// launch (blockIdx.z), any strides (rows of W and X contiguous).  Replaces the cuBLAS route of ggml_cuda_mul_mat (ggml-cuda.cu:2001-2084, ggml_cuda_op_mul_mat_cublas).
template <typename WT, typename XT>
__global__ void __launch_bounds__(256) k_mm_simt(const MmvfArgs A) {
    constexpr int BM = 64, BN = 64, BK = 32;
    __shared__ float Ws[BK][BM + 1], Xs[BK][BN + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t) blockIdx.x * BM, n0 = (int64_t) blockIdx.y * BN;
    const int64_t i2 = blockIdx.z % A.ne2, i3 = blockIdx.z / A.ne2;
    const char * wb = A.w + (i2 / A.r2) * A.w_nb2 + (i3 / A.r3) * A.w_nb3;
    const char * xb = A.x + i2 * A.x_nb2 + i3 * A.x_nb3;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int64_t k0 = 0; k0 < A.k; k0 += BK) {
        // a warp loads 32 consecutive k of one row (coalesced) and stores them down a column of the transposed tile (stride BM + 1: conflict-free)
#pragma unroll
        for (int e = 0; e < BM * BK / 256; ++e) {
            const int idx = e * 256 + tid, r = idx >> 5, kk = idx & 31;
            const bool kin = k0 + kk < A.k;
            Ws[kk][r] = kin && m0 + r < A.m ? w2f<WT>(((const WT *) (wb + (m0 + r) * A.w_nb1))[k0 + kk]) : 0.0f;
            Xs[kk][r] = kin && n0 + r < A.n ? round_like<WT>(w2f<XT>(((const XT *) (xb + (n0 + r) * A.x_nb1))[k0 + kk])) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = Ws[kk][tx + 16 * i]; b[i] = Xs[kk][ty + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    char * yb = A.y + i2 * A.y_nb2 + i3 * A.y_nb3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t c = n0 + ty + 16 * j;
        if (c >= A.n) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t r = m0 + tx + 16 * i;
            if (r < A.m) *(float *) (yb + c * A.y_nb1 + r * 4) = acc[i][j];
        }
    }
}

template <typename WT>
static int run_mm_simt(const MmvfArgs & A, bool x_f16, cudaStream_t st) {
    const int64_t nz = A.ne2 * A.ne3, gy = (A.n + 63) / 64;
    if (nz > 65535 || gy > 65535) return B200_ERR_UNSUPPORTED;
    const dim3 grid((unsigned) ((A.m + 63) / 64), (unsigned) gy, (unsigned) nz);
    if (x_f16) k_mm_simt<WT, __half><<<grid, 256, 0, st>>>(A);
    else       k_mm_simt<WT, float><<<grid, 256, 0, st>>>(A);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

} // namespace b200

using namespace b200;

extern "C" int b200_mul_mat_supported(const b200_tensor * w, const b200_tensor * x, const b200_tensor * dst) {
    if (!w || !x || !dst) return 0;
    const int t = w->type;
    if (!(is_quant(t) || t == B200_F32 || t == B200_F16 || t == B200_BF16)) return 0;
    // F16 activations: the im2col matrix x kernel product of ggml_conv_1d / conv_2d (F16 x F16 only, as the CPU backend: k_mm_simt reads them in place)
    const bool x16 = x->type == B200_F16 && t == B200_F16;          // (the only F16-activation product the CPU backend has: src1 type == vec_dot_type)
    if ((x->type != B200_F32 && !x16) || dst->type != B200_F32) return 0;
    const int64_t k = w->ne[0];
    if (k != x->ne[0] || k % blck_size(t) != 0) return 0;
    if (dst->ne[0] != w->ne[1] || dst->ne[1] != x->ne[1] || dst->ne[2] != x->ne[2] || dst->ne[3] != x->ne[3]) return 0;
    if (w->ne[2] == 0 || w->ne[3] == 0 || x->ne[2] % w->ne[2] != 0 || x->ne[3] % w->ne[3] != 0) return 0;
    if (w->nb[0] != type_size(t) || x->nb[0] != (x16 ? 2 : 4) || dst->nb[0] != 4) return 0;       // rows contiguous
    if (x16 && (x->ne[2] * x->ne[3] > 65535 || (x->ne[1] + 63) / 64 > 65535)) return 0;
    if (is_quant(t)) {
        if ((uintptr_t) x->data % 16 || x->nb[1] % 16 || x->nb[2] % 16 || x->nb[3] % 16) return 0;
        if (dst->nb[1] % 4) return 0;
        if (w->layout == B200_LAYOUT_PLANAR && (w->ne[2] != 1 || w->ne[3] != 1)) return 0;
    }
    return 1;
}

extern "C" size_t b200_mul_mat_scratch_bytes(const b200_tensor * w, const b200_tensor * x) {
    if (!is_quant(w->type)) return x->type == B200_F32 && mm_f16_tc_supported(w->type, w->ne[0], x->ne[1], w->data, w->nb[1]) && (uintptr_t) x->data % 16 == 0 && x->nb[1] % 16 == 0 ?
                                   mmq_tc_scratch_bytes(w->ne[0], x->ne[1]) : 0;
    const size_t act = (size_t) act_layout(w->type, w->ne[0]).bytes * (size_t) (x->ne[1] * x->ne[2] * x->ne[3]);
    // the tensor-core path keeps ONE batch slice of F16 activation tiles in the same scratch
    const size_t tc = mmq_tc_supported(w->type, w->layout, w->ne[0], x->ne[1], w->data, w->nb[1]) ? mmq_tc_scratch_bytes(w->ne[0], x->ne[1]) : 0;
    return act > tc ? act : tc;
}

extern "C" int b200_mul_mat(const b200_tensor * w, const b200_tensor * x, const b200_tensor * dst, void * scratch,
                            size_t scratch_bytes, void * stream) {
    return b200_mul_mat_ex(w, x, dst, scratch, scratch_bytes, 0, stream);
}

// dst = W . x + residual: the ADD that follows wo / ffn_down in every llama-family layer rides in the tensor-core GEMM's epilogue (quantised weights, 2-D
// operands, no split-K); every other case runs the MUL_MAT into dst and then the ADD kernel — which needs a residual that is NOT dst (the MUL_MAT would overwrite
// it): that combination returns B200_ERR_UNSUPPORTED before anything is launched, and the caller keeps its two separate ops.
extern "C" int b200_mul_mat_add(const b200_tensor * w, const b200_tensor * x, const b200_tensor * residual, const b200_tensor * dst, void * scratch,
                                size_t scratch_bytes, int flags, void * stream) {
    if (!residual || !b200_mul_mat_supported(w, x, dst) || x->type != B200_F32) return B200_ERR_UNSUPPORTED;
    if (residual->type != B200_F32 || residual->ne[0] != dst->ne[0] || residual->ne[1] != dst->ne[1] || residual->ne[2] != dst->ne[2] || residual->ne[3] != dst->ne[3]) return B200_ERR_UNSUPPORTED;
    const int t = w->type;
    const int64_t k = w->ne[0], m = w->ne[1], n = x->ne[1];
    const bool fusable = is_quant(t) && !tc_disabled() && mmq_tc_supported(t, w->layout, k, n, w->data, w->nb[1]) && x->ne[2] * x->ne[3] == 1 && w->ne[2] * w->ne[3] == 1 &&
                         residual->nb[0] == 4 && residual->nb[1] == dst->nb[1] && dst->nb[0] == 4 && scratch && (uintptr_t) scratch % 16 == 0 &&
                         scratch_bytes >= b200_mul_mat_scratch_bytes(w, x) && m > 0 && n > 0 && mmq_tc_fuses_resid(m, k, n);
    if (fusable) {
        bool fused = false;
        const int rc = mmq_tc(w->data, t, m, k, (const float *) x->data, x->nb[1] / 4, n, (float *) dst->data, dst->nb[1] / 4, scratch, (flags & B200_MM_REUSE_ACT) != 0,
                              (cudaStream_t) stream, (const float *) residual->data, &fused);
        return rc ? rc : fused ? B200_OK : B200_ERR_ARG;                    // mmq_tc_fuses_resid and mmq_tc share tc_splitk: not fused here would be a bug
    }
    // decode (ONE column, 2-D weight): the residual rides in the matvec's epilogue — each thread reads residual[i] before it writes y[i], so residual may be dst
    if (n == 1 && x->ne[2] * x->ne[3] == 1 && w->ne[2] * w->ne[3] == 1 && residual->nb[0] == 4 && dst->nb[0] == 4 && m > 0) {
        if (is_quant(t) && scratch && (uintptr_t) scratch % 16 == 0 && scratch_bytes >= (size_t) act_layout(t, k).bytes && (uintptr_t) x->data % 16 == 0) {
            int rc = (flags & B200_MM_REUSE_ACT) ? B200_OK : b200_quantize_act(t, (const float *) x->data, x->nb[1] / 4, scratch, k, 1, stream);
            if (rc) return rc;
            b200_matvec_job job = { w->data, t, w->layout, m, w->nb[1], (float *) dst->data, (const float *) residual->data };
            rc = b200_matvec_q(&job, 1, scratch, k, stream);
            if (rc != B200_ERR_UNSUPPORTED) return rc;
        } else if (t == B200_F16) {
            MmvfArgs A = { (const char *) w->data, (const char *) x->data, (char *) dst->data, m, k, n, w->nb[1], w->nb[2], w->nb[3], x->nb[1], x->nb[2], x->nb[3],
                           dst->nb[1], dst->nb[2], dst->nb[3], 1, 1, 1, 1 };
            if (mmvf16_stream_ok(A)) return run_mmvf16_stream(A, nullptr, (const float *) residual->data, (cudaStream_t) stream);
        }
    }
    const char * r0 = (const char *) residual->data, * d0 = (const char *) dst->data;
    const int64_t rbytes = residual->nb[3] * residual->ne[3], dbytes = dst->nb[3] * dst->ne[3];
    if (r0 < d0 + dbytes && d0 < r0 + rbytes) return B200_ERR_UNSUPPORTED;   // overlapping residual and dst
    const int rc = b200_mul_mat_ex(w, x, dst, scratch, scratch_bytes, flags, stream);
    return rc ? rc : b200_binary(B200_ADD, dst, residual, dst, stream);
}

// dst[i] = W[i] . x for 2 or 3 weight matrices over the SAME activations (q / k / v): ONE tensor-core launch over the concatenated m-tiles when every W[i] is a K-quant
// on the tcgen05 path (the 1024-row wk / wv alone fill 64 of 148 SMs; merged with wq the launch has 48 x n/256 tiles), otherwise one MUL_MAT after the other with the
// activation tiles shared.  scratch: the largest b200_mul_mat_scratch_bytes of the group.
// dst = silu(Wg . x) * (Wu . x) for ONE activation column (the gate / up / SWIGLU triple of a decode graph, three adjacent nodes): one launch — the q8 activation
// record is built once and b200_matvec_q_swiglu streams both matrices (quantised weights), or k_mmvf16_stream<GLU> (F16 weights).  B200_ERR_UNSUPPORTED for everything
// else (other GLU ops, several columns, batch dims): the caller keeps its three ops.  Replaces mul_mat_vec_q x 2 + unary_gated_op_kernel (mmvq.cu:139, unary.cu:208-228).
extern "C" int b200_mul_mat_glu(int glu_op, const b200_tensor * w_gate, const b200_tensor * w_up, const b200_tensor * x, const b200_tensor * dst, void * scratch,
                                size_t scratch_bytes, int flags, void * stream) {
    if (!w_gate || !w_up || !x || !dst) return B200_ERR_ARG;
    if (glu_op != B200_GLU_SWIGLU || x->type != B200_F32 || !b200_mul_mat_supported(w_gate, x, dst) || !b200_mul_mat_supported(w_up, x, dst)) return B200_ERR_UNSUPPORTED;
    const int t = w_gate->type;
    const int64_t k = w_gate->ne[0], m = w_gate->ne[1];
    if (w_up->type != t || w_up->ne[1] != m || w_up->layout != w_gate->layout || w_up->nb[1] != w_gate->nb[1] || x->ne[1] != 1 || x->ne[2] * x->ne[3] != 1 ||
        w_gate->ne[2] * w_gate->ne[3] != 1 || w_up->ne[2] * w_up->ne[3] != 1 || dst->nb[0] != 4 || m <= 0) return B200_ERR_UNSUPPORTED;
    if (is_quant(t)) {
        if (!scratch || (uintptr_t) scratch % 16 || scratch_bytes < (size_t) act_layout(t, k).bytes || (uintptr_t) x->data % 16) return B200_ERR_UNSUPPORTED;
        const int rc = (flags & B200_MM_REUSE_ACT) ? B200_OK : b200_quantize_act(t, (const float *) x->data, x->nb[1] / 4, scratch, k, 1, stream);
        if (rc) return rc;
        const b200_matvec_job g = { w_gate->data, t, w_gate->layout, m, w_gate->nb[1], nullptr, nullptr }, u = { w_up->data, t, w_up->layout, m, w_up->nb[1], nullptr, nullptr };
        return b200_matvec_q_swiglu(&g, &u, (float *) dst->data, scratch, k, stream);
    }
    if (t != B200_F16) return B200_ERR_UNSUPPORTED;
    MmvfArgs A = { (const char *) w_gate->data, (const char *) x->data, (char *) dst->data, m, k, 1, w_gate->nb[1], w_gate->nb[2], w_gate->nb[3], x->nb[1], x->nb[2], x->nb[3],
                   dst->nb[1], dst->nb[2], dst->nb[3], 1, 1, 1, 1 };
    if (!mmvf16_stream_ok(A) || (uintptr_t) w_up->data % 16) return B200_ERR_UNSUPPORTED;
    return run_mmvf16_stream(A, w_up->data, nullptr, (cudaStream_t) stream);
}

// B200_NO_MULTI=1 keeps every MUL_MAT of a group in its own launch (read per call: the tests toggle it to compare the merged launch with the separate ones)
static bool multi_disabled() { const char * e = getenv("B200_NO_MULTI"); return e && atoi(e) != 0; }
// would b200_mul_mat_multi run this group as ONE launch?  (a caller that has to stage the later results elsewhere — the ggml plugin — only does so when it pays)
extern "C" int b200_mul_mat_multi_merges(int n_mat, const b200_tensor * const * w, const b200_tensor * x) {
    if (n_mat < 2 || n_mat > 3 || !w || !x || tc_disabled() || multi_disabled() || x->ne[2] * x->ne[3] != 1 || x->type != B200_F32) return 0;
    const int64_t k = x->ne[0], n = x->ne[1];
    int type[3]; int64_t m[3];
    for (int i = 0; i < n_mat; ++i) {
        if (!w[i] || w[i]->ne[0] != k || w[i]->ne[2] * w[i]->ne[3] != 1 || !is_quant(w[i]->type) || !mmq_tc_supported(w[i]->type, w[i]->layout, k, n, w[i]->data, w[i]->nb[1])) return 0;
        type[i] = w[i]->type; m[i] = w[i]->ne[1];
    }
    return mmq_tc_multi_ok(n_mat, type, m, k, n) ? 1 : 0;
}

extern "C" int b200_mul_mat_multi(int n_mat, const b200_tensor * const * w, const b200_tensor * x, const b200_tensor * const * dst, void * scratch, size_t scratch_bytes,
                                  int flags, void * stream) {
    if (n_mat < 1 || n_mat > 3 || !w || !x || !dst) return B200_ERR_ARG;
    for (int i = 0; i < n_mat; ++i) if (!b200_mul_mat_supported(w[i], x, dst[i])) return B200_ERR_UNSUPPORTED;
    if (x->type != B200_F32) return B200_ERR_UNSUPPORTED;
    const int64_t k = x->ne[0], n = x->ne[1];
    bool merge = n_mat > 1 && !tc_disabled() && !multi_disabled() && x->ne[2] * x->ne[3] == 1 && scratch && (uintptr_t) scratch % 16 == 0 && n > 0;
    int type[3]; int64_t m[3], ld[3]; const void * wp[3]; float * dp[3];
    for (int i = 0; i < n_mat && merge; ++i) {
        type[i] = w[i]->type; m[i] = w[i]->ne[1]; ld[i] = dst[i]->nb[1] / 4; wp[i] = w[i]->data; dp[i] = (float *) dst[i]->data;
        merge = is_quant(type[i]) && mmq_tc_supported(type[i], w[i]->layout, k, n, w[i]->data, w[i]->nb[1]) && w[i]->ne[2] * w[i]->ne[3] == 1 && dst[i]->nb[0] == 4 &&
                scratch_bytes >= b200_mul_mat_scratch_bytes(w[i], x);
    }
    if (merge && mmq_tc_multi_ok(n_mat, type, m, k, n))
        return mmq_tc_multi(n_mat, wp, type, m, k, (const float *) x->data, x->nb[1] / 4, n, dp, ld, scratch, (flags & B200_MM_REUSE_ACT) != 0, (cudaStream_t) stream, nullptr, nullptr);
    for (int i = 0; i < n_mat; ++i) {
        const int rc = b200_mul_mat_ex(w[i], x, dst[i], scratch, scratch_bytes, i == 0 ? flags : (flags | B200_MM_REUSE_ACT), stream);
        if (rc) return rc;
    }
    return B200_OK;
}

extern "C" int b200_mul_mat_ex(const b200_tensor * w, const b200_tensor * x, const b200_tensor * dst, void * scratch,
                               size_t scratch_bytes, int flags, void * stream) {
    if (!b200_mul_mat_supported(w, x, dst)) return B200_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t) stream;
    const int t = w->type;
    const int64_t k = w->ne[0], m = w->ne[1], n = x->ne[1];
    if (m == 0 || n == 0 || dst->ne[2] * dst->ne[3] == 0) return B200_OK;
    const int64_t r2 = x->ne[2] / w->ne[2], r3 = x->ne[3] / w->ne[3];
    if (!is_quant(t)) {
        // F16 weights, more than 8 columns: TMA-fed tcgen05 GEMM (mmq_tc.cu k_mm_f16_tc), one launch per batch slice
        const bool x16 = x->type == B200_F16;
        // many SMALL batch slices (attention products per head: Whisper's K.Q^T is 16 heads x [T x 50 x 64]) are launch-bound on the per-slice tensor-core route
        // (tile conversion + GEMM launch per slice): below ~50 MFLOP per slice the one-launch SIMT GEMM is faster
        const bool small_slices = x->ne[2] * x->ne[3] > 1 && 2.0 * (double) m * (double) n * (double) k < 5e7 && !simt_disabled() && x->ne[2] * x->ne[3] <= 65535;
        if (!x16 && !small_slices && !tc_disabled() && mm_f16_tc_supported(t, k, n, w->data, w->nb[1]) && scratch && (uintptr_t) scratch % 16 == 0 &&
            scratch_bytes >= mmq_tc_scratch_bytes(k, n) && (uintptr_t) x->data % 16 == 0 && x->nb[1] % 16 == 0 && x->nb[2] % 16 == 0 && x->nb[3] % 16 == 0 &&
            w->nb[2] % 16 == 0 && w->nb[3] % 16 == 0 && dst->nb[1] % 4 == 0) {
            for (int64_t i3 = 0; i3 < x->ne[3]; ++i3) for (int64_t i2 = 0; i2 < x->ne[2]; ++i2) {
                const float * xs = (const float *) ((const char *) x->data + i2 * x->nb[2] + i3 * x->nb[3]);
                const char * wb = (const char *) w->data + (i2 / r2) * w->nb[2] + (i3 / r3) * w->nb[3];
                float * yb = (float *) ((char *) dst->data + i2 * dst->nb[2] + i3 * dst->nb[3]);
                const int rc = mm_f16_tc(wb, m, k, xs, x->nb[1] / 4, n, yb, dst->nb[1] / 4, scratch, (flags & B200_MM_REUSE_ACT) && x->ne[2] * x->ne[3] == 1, st);
                if (rc) return rc;
            }
            return B200_OK;
        }
        MmvfArgs A = { (const char *) w->data, (const char *) x->data, (char *) dst->data, m, k, n,
                       w->nb[1], w->nb[2], w->nb[3], x->nb[1], x->nb[2], x->nb[3], dst->nb[1], dst->nb[2], dst->nb[3],
                       dst->ne[2], dst->ne[3], r2, r3 };
        // more than 8 columns (or F16 activations): the tiled GEMM, one launch over all batch slices; up to 8 columns: the warp-per-row matvec
        if (x16 || (n > 8 && !simt_disabled() && dst->ne[2] * dst->ne[3] <= 65535 && (n + 63) / 64 <= 65535)) {
            if (t == B200_F32)  return run_mm_simt<float>(A, x16, st);
            if (t == B200_F16)  return run_mm_simt<__half>(A, x16, st);
            return run_mm_simt<__nv_bfloat16>(A, x16, st);
        }
        if (t == B200_F32)  return run_mmvf<float>(A, st);
        if (t == B200_F16)  return run_mmvf<__half>(A, st);
        return run_mmvf<__nv_bfloat16>(A, st);
    }
    if (scratch_bytes < b200_mul_mat_scratch_bytes(w, x) || ((uintptr_t) scratch % 16)) return B200_ERR_ARG;
    const int64_t act_b = act_layout(t, k).bytes;
    const bool tc = !tc_disabled() && mmq_tc_supported(t, w->layout, k, n, w->data, w->nb[1]);
    for (int64_t i3 = 0; i3 < x->ne[3]; ++i3) for (int64_t i2 = 0; i2 < x->ne[2]; ++i2) {
        const float * xs = (const float *) ((const char *) x->data + i2 * x->nb[2] + i3 * x->nb[3]);
        if (tc) {       // prefill / batched: dequant tiles -> tcgen05.mma (mmq_tc.cu)
            const char * wb = (const char *) w->data + (i2 / r2) * w->nb[2] + (i3 / r3) * w->nb[3];
            float * yb = (float *) ((char *) dst->data + i2 * dst->nb[2] + i3 * dst->nb[3]);
            const int rc = mmq_tc(wb, t, m, k, xs, x->nb[1] / 4, n, yb, dst->nb[1] / 4, scratch, (flags & B200_MM_REUSE_ACT) && x->ne[2] * x->ne[3] == 1, st, nullptr, nullptr);
            if (rc) return rc;
            continue;
        }
        uint8_t * act = (uint8_t *) scratch + (i3 * x->ne[2] + i2) * n * act_b;
        // B200_MM_REUSE_ACT on the matvec path: the scratch still holds the q8 records of THIS x from the previous MUL_MAT of the same record class (q / k / v)
        int rc = (flags & B200_MM_REUSE_ACT) && x->ne[2] * x->ne[3] == 1 ? B200_OK : b200_quantize_act(t, xs, x->nb[1] / 4, act, k, n, stream);
        if (rc) return rc;
        const char * wb = (const char *) w->data + (i2 / r2) * w->nb[2] + (i3 / r3) * w->nb[3];
        float * yb = (float *) ((char *) dst->data + i2 * dst->nb[2] + i3 * dst->nb[3]);
        for (int64_t c0 = 0; c0 < n; c0 += 8) {
            const int nc = (int) (n - c0 < 8 ? n - c0 : 8);
            rc = matvec_q_cols(wb, t, w->layout, m, w->nb[1], act + c0 * act_b, k, nc, yb + c0 * (dst->nb[1] / 4), dst->nb[1] / 4, st);
            if (rc) return rc;
        }
    }
    return B200_OK;
}
