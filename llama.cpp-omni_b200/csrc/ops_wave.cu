// ops_wave.cu — the op set of the Token2Wav graphs (flow-matching CFM + HiFiGAN vocoder) that the LLM / encoder paths do not
// already need.  tools/omni/token2wav/token2wav-impl.cpp runs its graphs with a direct graph_compute on the first GPU-type device (no scheduler, no
// CPU fallback: :1905-1916), so every op it builds must exist here:
//   CONCAT              replaces concat_f32_dim0/1/2 + concat_f32_non_cont   ggml-cuda/concat.cu            CPU ggml-cpu/ops.cpp:1839-2040
//   REPEAT              replaces k_repeat (bin_bcast op_repeat)              ggml-cuda/binbcast.cu          CPU ops.cpp:1637-1700
//   ARANGE              replaces arange_f32                                  ggml-cuda/arange.cu            CPU ops.cpp:7762-7785
//   SUM_ROWS            replaces k_sum_rows_f32 / reduce_rows_f32            ggml-cuda/sumrows.cu           CPU ops.cpp:1399-1430 (f64 accumulator)
//   PAD                 replaces pad_f32                                     ggml-cuda/pad.cu               CPU ops.cpp:7592-7640
//   PAD_REFLECT_1D      replaces pad_reflect_1d_kernel_f32                   ggml-cuda/pad_reflect_1d.cu    CPU ops.cpp:7664-7692
//   CONV_TRANSPOSE_1D   replaces conv_transpose_1d_kernel                    ggml-cuda/conv-transpose-1d.cu CPU ops.cpp:5952-6130
// (LEAKY_RELU / ELU / SIN / COS / LOG / STEP / SGN / CLAMP / HARDSWISH / HARDSIGMOID are b200_unary_param in ops_misc.cu.)
// All of them are index-remapping copies or short reductions: HBM/L2-bound, one launch each, grid-stride, strides taken from nb[] so views work.
#include "common.cuh"

namespace b200 {

struct W4 { char * data; int type; int64_t ne[4]; int64_t nb[4]; };
static inline W4 w4(const b200_tensor * t) {
    W4 r; r.data = (char *) t->data; r.type = t->type;
    for (int i = 0; i < 4; ++i) { r.ne[i] = t->ne[i]; r.nb[i] = t->nb[i]; }
    return r;
}
static inline int64_t wn(const b200_tensor * t) { return t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3]; }
static inline unsigned wgrid(int64_t items) {
    int64_t g = (items + 255) / 256;
    const int64_t cap = (int64_t) sm_count() * 16;
    return (unsigned) (g > cap ? cap : g < 1 ? 1 : g);
}
static inline bool copyable(int t) { return t == B200_F32 || t == B200_F16 || t == B200_BF16 || t == B200_I32; }

__device__ __forceinline__ void copy_elem(char * d, const char * s, int ts) {
    if (ts == 4) *(uint32_t *) d = *(const uint32_t *) s; else *(uint16_t *) d = *(const uint16_t *) s;
}
__device__ __forceinline__ void split4(int64_t g, const int64_t * ne, int64_t & i0, int64_t & i1, int64_t & i2, int64_t & i3) {
    i0 = g % ne[0]; g /= ne[0]; i1 = g % ne[1]; g /= ne[1]; i2 = g % ne[2]; i3 = g / ne[2];
}

// ---- CONCAT: dst = [a ; b] along `dim`
struct ConcatArgs { W4 a, b, dst; int dim, ts; };
__global__ void __launch_bounds__(256) k_concat(const ConcatArgs A, int64_t total) {
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t) gridDim.x * blockDim.x) {
        int64_t i[4]; split4(g, A.dst.ne, i[0], i[1], i[2], i[3]);
        char * d = A.dst.data + i[0] * A.dst.nb[0] + i[1] * A.dst.nb[1] + i[2] * A.dst.nb[2] + i[3] * A.dst.nb[3];
        const bool first = i[A.dim] < A.a.ne[A.dim];
        const W4 & S = first ? A.a : A.b;
        if (!first) i[A.dim] -= A.a.ne[A.dim];
        copy_elem(d, S.data + i[0] * S.nb[0] + i[1] * S.nb[1] + i[2] * S.nb[2] + i[3] * S.nb[3], A.ts);
    }
}

// ---- REPEAT: dst[i] = src[i mod src.ne]
struct RepeatArgs { W4 src, dst; int ts; };
__global__ void __launch_bounds__(256) k_repeat(const RepeatArgs A, int64_t total) {
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t) gridDim.x * blockDim.x) {
        int64_t i0, i1, i2, i3; split4(g, A.dst.ne, i0, i1, i2, i3);
        copy_elem(A.dst.data + i0 * A.dst.nb[0] + i1 * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3],
                  A.src.data + (i0 % A.src.ne[0]) * A.src.nb[0] + (i1 % A.src.ne[1]) * A.src.nb[1] + (i2 % A.src.ne[2]) * A.src.nb[2] + (i3 % A.src.ne[3]) * A.src.nb[3], A.ts);
    }
}

// ---- ARANGE: dst[i] = start + step * i  (the reference's expression, one multiply and one add in f32)
__global__ void __launch_bounds__(256) k_arange(float * __restrict__ d, int64_t n, float start, float step) {
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (int64_t) gridDim.x * blockDim.x) d[g] = __fadd_rn(start, __fmul_rn(step, (float) g));
}

// ---- SUM_ROWS: one warp per row, f64 partial sums per lane (the reference accumulates the row in ggml_float = double), fixed shuffle tree
struct SumArgs { W4 x, dst; };
__global__ void __launch_bounds__(256) k_sum_rows(const SumArgs A, int64_t rows) {
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t) blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t) gridDim.x * 8) {
        const int64_t i1 = r % A.x.ne[1], i2 = (r / A.x.ne[1]) % A.x.ne[2], i3 = r / (A.x.ne[1] * A.x.ne[2]);
        const float * x = (const float *) (A.x.data + i1 * A.x.nb[1] + i2 * A.x.nb[2] + i3 * A.x.nb[3]);
        double s = 0.0;
        for (int64_t i = lane; i < A.x.ne[0]; i += 32) s += (double) x[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) *(float *) (A.dst.data + i1 * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3]) = (float) s;
    }
}

// ---- PAD: zero padding left / right of every dimension (op_params lp0, rp0, ..., lp3, rp3)
struct PadArgs { W4 x, dst; int lp[4]; };
__global__ void __launch_bounds__(256) k_pad(const PadArgs A, int64_t total) {
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t) gridDim.x * blockDim.x) {
        int64_t i[4]; split4(g, A.dst.ne, i[0], i[1], i[2], i[3]);
        bool in = true;
        const char * s = A.x.data;
#pragma unroll
        for (int d = 0; d < 4; ++d) { const int64_t j = i[d] - A.lp[d]; in = in && j >= 0 && j < A.x.ne[d]; s += j * A.x.nb[d]; }
        *(float *) (A.dst.data + i[0] * A.dst.nb[0] + i[1] * A.dst.nb[1] + i[2] * A.dst.nb[2] + i[3] * A.dst.nb[3]) = in ? *(const float *) s : 0.0f;
    }
}

// ---- PAD_REFLECT_1D: dst[p0 + i] = x[i]; dst[p0 - i] = x[i] (i = 1..p0); dst[p0 + n - 1 + i] = x[n - 1 - i] (i = 1..p1)
struct ReflArgs { W4 x, dst; int p0, p1; };
__global__ void __launch_bounds__(256) k_pad_reflect_1d(const ReflArgs A, int64_t total) {
    const int64_t n = A.x.ne[0];
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t) gridDim.x * blockDim.x) {
        int64_t i0, i1, i2, i3; split4(g, A.dst.ne, i0, i1, i2, i3);
        int64_t j = i0 - A.p0;
        if (j < 0) j = -j; else if (j >= n) j = 2 * (n - 1) - j;
        *(float *) (A.dst.data + i0 * A.dst.nb[0] + i1 * A.dst.nb[1] + i2 * A.dst.nb[2] + i3 * A.dst.nb[3]) =
            *(const float *) (A.x.data + j * A.x.nb[0] + i1 * A.x.nb[1] + i2 * A.x.nb[2] + i3 * A.x.nb[3]);
    }
}

// ---- CONV_TRANSPOSE_1D (p0 = 0, d0 = 1: all the reference implements): kernel [K, Cout, Cin] F32 / F16, x [L, Cin] F32, dst [(L-1)*s0 + K, Cout] F32
//      dst[o, co] = sum over (l, k) with l*s0 + k == o of sum_ci x[l, ci] * w[k, co, ci]: a gather per output element, so no zero-fill pass and no atomics;
//      at most ceil(K / s0) taps contribute.  One THREAD per output, consecutive threads = consecutive o: a warp's x reads (x[l, ci], l = (o - k) / s0) fall into
//      one 32-byte sector and its w reads (w[k, co, ci], k = o - l*s0) into the K contiguous taps of one (co, ci) — a warp-per-output version with the lanes over Cin
//      had every lane on its own sector (x and w are both strided in ci) and ran 40x slower on the HiFiGAN upsampling shapes.
struct CtArgs { W4 w, x, dst; int s0; };
template <typename TW>
__global__ void __launch_bounds__(256) k_conv_transpose_1d(const CtArgs A, int64_t total) {
    const int64_t K = A.w.ne[0], Cin = A.w.ne[2], L = A.x.ne[0], OL = A.dst.ne[0];
    for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t) gridDim.x * blockDim.x) {
        const int64_t o = g % OL, co = g / OL;
        const int64_t l_lo = o - (K - 1) <= 0 ? 0 : (o - (K - 1) + A.s0 - 1) / A.s0;
        int64_t l_hi = o / A.s0; if (l_hi > L - 1) l_hi = L - 1;
        const char * wco = A.w.data + co * A.w.nb[1];
        float acc = 0.0f;
        const int nt = (int) (l_hi - l_lo + 1);
        if (nt <= 2) {                                                         // the vocoder's case (K = 2 * stride): the taps are fixed per output, only two pointers walk Cin
            const char * x0 = A.x.data + l_lo * 4, * w0 = wco + (o - l_lo * A.s0) * (int64_t) sizeof(TW);
            const int64_t xs = A.x.nb[1], ws = A.w.nb[2], wd = -(int64_t) A.s0 * (int64_t) sizeof(TW);
            float acc1 = 0.0f;
            if (nt == 2) {
#pragma unroll 4
                for (int64_t ci = 0; ci < Cin; ++ci, x0 += xs, w0 += ws) {
                    acc  = fmaf(*(const float *) x0, (float) *(const TW *) w0, acc);
                    acc1 = fmaf(*(const float *) (x0 + 4), (float) *(const TW *) (w0 + wd), acc1);
                }
            } else if (nt == 1) {
#pragma unroll 4
                for (int64_t ci = 0; ci < Cin; ++ci, x0 += xs, w0 += ws) acc = fmaf(*(const float *) x0, (float) *(const TW *) w0, acc);
            }
            acc += acc1;
        } else {
            for (int64_t ci = 0; ci < Cin; ++ci) {
                const float * xr = (const float *) (A.x.data + ci * A.x.nb[1]);
                const TW * wr = (const TW *) (wco + ci * A.w.nb[2]);
                for (int64_t l = l_lo; l <= l_hi; ++l) acc = fmaf(xr[l], (float) wr[o - l * A.s0], acc);
            }
        }
        *(float *) (A.dst.data + o * A.dst.nb[0] + co * A.dst.nb[1]) = acc;
    }
}

} // namespace b200
using namespace b200;

extern "C" int b200_concat(const b200_tensor * a, const b200_tensor * b, const b200_tensor * dst, int dim, void * stream) {
    if (!a || !b || !dst) return B200_ERR_ARG;
    if (dim < 0 || dim > 3 || a->type != b->type || a->type != dst->type || !copyable(a->type)) return B200_ERR_UNSUPPORTED;
    for (int d = 0; d < 4; ++d) {
        if (d == dim) { if (dst->ne[d] != a->ne[d] + b->ne[d]) return B200_ERR_UNSUPPORTED; }
        else if (a->ne[d] != b->ne[d] || dst->ne[d] != a->ne[d]) return B200_ERR_UNSUPPORTED;
    }
    const int64_t total = wn(dst);
    if (total == 0) return B200_OK;
    ConcatArgs A = { w4(a), w4(b), w4(dst), dim, type_size(a->type) };
    k_concat<<<wgrid(total), 256, 0, (cudaStream_t) stream>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_repeat(const b200_tensor * src, const b200_tensor * dst, void * stream) {
    if (!src || !dst) return B200_ERR_ARG;
    if (src->type != dst->type || !copyable(src->type)) return B200_ERR_UNSUPPORTED;
    for (int d = 0; d < 4; ++d) if (src->ne[d] == 0 || dst->ne[d] % src->ne[d]) return wn(dst) == 0 ? B200_OK : B200_ERR_UNSUPPORTED;
    const int64_t total = wn(dst);
    if (total == 0) return B200_OK;
    RepeatArgs A = { w4(src), w4(dst), type_size(src->type) };
    k_repeat<<<wgrid(total), 256, 0, (cudaStream_t) stream>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_arange(const b200_tensor * dst, float start, float step, void * stream) {
    if (!dst) return B200_ERR_ARG;
    if (dst->type != B200_F32 || dst->nb[0] != 4 || dst->ne[1] * dst->ne[2] * dst->ne[3] != 1) return B200_ERR_UNSUPPORTED;
    if (dst->ne[0] == 0) return B200_OK;
    k_arange<<<wgrid(dst->ne[0]), 256, 0, (cudaStream_t) stream>>>((float *) dst->data, dst->ne[0], start, step);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_sum_rows(const b200_tensor * x, const b200_tensor * dst, void * stream) {
    if (!x || !dst) return B200_ERR_ARG;
    if (x->type != B200_F32 || dst->type != B200_F32 || x->nb[0] != 4 || dst->ne[0] != 1 || dst->ne[1] != x->ne[1] || dst->ne[2] != x->ne[2] || dst->ne[3] != x->ne[3])
        return B200_ERR_UNSUPPORTED;
    const int64_t rows = x->ne[1] * x->ne[2] * x->ne[3];
    if (rows == 0) return B200_OK;
    SumArgs A = { w4(x), w4(dst) };
    k_sum_rows<<<wgrid(rows * 32), 256, 0, (cudaStream_t) stream>>>(A, rows);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_pad(const b200_tensor * x, const b200_tensor * dst, const int32_t * lp_rp8, void * stream) {
    if (!x || !dst || !lp_rp8) return B200_ERR_ARG;
    if (x->type != B200_F32 || dst->type != B200_F32) return B200_ERR_UNSUPPORTED;
    PadArgs A; A.x = w4(x); A.dst = w4(dst);
    for (int d = 0; d < 4; ++d) {
        A.lp[d] = lp_rp8[2 * d];
        if (lp_rp8[2 * d] < 0 || lp_rp8[2 * d + 1] < 0 || dst->ne[d] != x->ne[d] + lp_rp8[2 * d] + lp_rp8[2 * d + 1]) return B200_ERR_UNSUPPORTED;
    }
    const int64_t total = wn(dst);
    if (total == 0) return B200_OK;
    k_pad<<<wgrid(total), 256, 0, (cudaStream_t) stream>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_pad_reflect_1d(const b200_tensor * x, const b200_tensor * dst, int p0, int p1, void * stream) {
    if (!x || !dst) return B200_ERR_ARG;
    if (x->type != B200_F32 || dst->type != B200_F32 || p0 < 0 || p1 < 0 || p0 >= x->ne[0] || p1 >= x->ne[0] || dst->ne[0] != x->ne[0] + p0 + p1 ||
        dst->ne[1] != x->ne[1] || dst->ne[2] != x->ne[2] || dst->ne[3] != x->ne[3]) return B200_ERR_UNSUPPORTED;
    const int64_t total = wn(dst);
    if (total == 0) return B200_OK;
    ReflArgs A = { w4(x), w4(dst), p0, p1 };
    k_pad_reflect_1d<<<wgrid(total), 256, 0, (cudaStream_t) stream>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_conv_transpose_1d(const b200_tensor * kernel, const b200_tensor * x, const b200_tensor * dst, int s0, void * stream) {
    if (!kernel || !x || !dst) return B200_ERR_ARG;
    if ((kernel->type != B200_F32 && kernel->type != B200_F16) || x->type != B200_F32 || dst->type != B200_F32 || s0 <= 0) return B200_ERR_UNSUPPORTED;
    if (kernel->ne[3] != 1 || x->ne[2] * x->ne[3] != 1 || dst->ne[2] * dst->ne[3] != 1 || kernel->ne[2] != x->ne[1] || dst->ne[1] != kernel->ne[1] ||
        dst->ne[0] != (x->ne[0] - 1) * s0 + kernel->ne[0]) return B200_ERR_UNSUPPORTED;
    const int64_t total = dst->ne[0] * dst->ne[1];
    if (total == 0) return B200_OK;
    CtArgs A = { w4(kernel), w4(x), w4(dst), s0 };
    if (kernel->nb[0] != type_size(kernel->type) || x->nb[0] != 4) return B200_ERR_UNSUPPORTED;
    if (kernel->type == B200_F32) k_conv_transpose_1d<float><<<wgrid(total), 256, 0, (cudaStream_t) stream>>>(A, total);
    else                          k_conv_transpose_1d<__half><<<wgrid(total), 256, 0, (cudaStream_t) stream>>>(A, total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
