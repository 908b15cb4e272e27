/* oracle/ops_port.c — TEST INFRASTRUCTURE (CPU oracle), never linked into the product.
 *
 * Plain-C restatement of the non-matmul ops on the Qwen3 decode path, following the reference CPU backend:
 *   RMS_NORM        ggml/src/ggml-cpu/ops.cpp:3517-3565   (double-precision sum of squares, scale = 1/sqrtf(mean+eps))
 *   ROPE            ops.cpp:5436-5720 (theta chain: theta *= theta_scale per pair; neox pairs (i, i+n_dims/2), norm pairs (2i, 2i+1));
 *                   YaRN helpers ggml.c:4122-4136
 *   GLU swiglu      ops.cpp:2934-2990 + vec.h ggml_vec_swiglu_f32: silu(g) * u, silu(x) = x / (1 + expf(-x))
 *   FLASH_ATTN_EXT  ops.cpp:7912-8148 (Q rounded to f16, f16 K.Q dot with f32 accumulate, ONLINE softmax, V accumulated in F16
 *                   when V is f16 — yes, the CPU oracle is lossy here)
 *   SET_ROWS        ops.cpp ggml_compute_forward_set_rows_f32 (F32 rows -> F16 rows at I64 indices)
 *   SOFT_MAX        ops.cpp ggml_compute_forward_soft_max_f32 (scale, mask, max-subtracted expf, 1/sum)
 * Pinned against oracle/_ref (single-op graphs on the reference CPU backend) by tests/test_oracle_pin.py.
 */
#include "blocks.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

void or_rms_norm(const float * x, float * y, int64_t ncols, int64_t nrows, float eps) {
    for (int64_t r = 0; r < nrows; ++r, x += ncols, y += ncols) {
        double ss = 0.0;
        for (int64_t i = 0; i < ncols; ++i) ss += (double) (x[i] * x[i]);
        const float mean  = (float) (ss / ncols);
        const float scale = 1.0f / sqrtf(mean + eps);
        for (int64_t i = 0; i < ncols; ++i) y[i] = x[i] * scale;
    }
}

void or_mul_rows(const float * x, const float * w, float * y, int64_t ncols, int64_t nrows) {   /* MUL with [ncols] broadcast */
    for (int64_t r = 0; r < nrows; ++r) for (int64_t i = 0; i < ncols; ++i) y[r*ncols + i] = x[r*ncols + i] * w[i];
}
void or_add(const float * a, const float * b, float * y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = a[i] + b[i]; }

/* ---- ROPE ------------------------------------------------------------------------------------ */
static float yarn_corr_dim(int n_dims, int n_ctx_orig, float n_rot, float base) {
    return n_dims * logf(n_ctx_orig / (n_rot * 2 * (float) M_PI)) / (2 * logf(base));
}
void or_rope_corr_dims(int n_dims, int n_ctx_orig, float freq_base, float beta_fast, float beta_slow, float dims[2]) {
    float lo = floorf(yarn_corr_dim(n_dims, n_ctx_orig, beta_fast, freq_base));
    float hi = ceilf (yarn_corr_dim(n_dims, n_ctx_orig, beta_slow, freq_base));
    dims[0] = lo < 0 ? 0 : lo;
    dims[1] = hi > n_dims - 1 ? n_dims - 1 : hi;
}
/* x, y: [n_tok][n_head][head_dim] contiguous;  pos: [n_tok];  mode: 0 = norm (adjacent pairs), 2 = neox (half-split pairs) */
void or_rope(const float * x, float * y, const int32_t * pos, const float * freq_factors,
             int64_t head_dim, int64_t n_head, int64_t n_tok, int n_dims, int mode, int n_ctx_orig,
             float freq_base, float freq_scale, float ext_factor, float attn_factor, float beta_fast, float beta_slow) {
    const float theta_scale = powf(freq_base, -2.0f / n_dims);
    float corr[2]; or_rope_corr_dims(n_dims, n_ctx_orig, freq_base, beta_fast, beta_slow, corr);
    float * cs = malloc(sizeof(float) * n_dims);
    for (int64_t t = 0; t < n_tok; ++t) {
        float theta = (float) pos[t];
        for (int i0 = 0; i0 < n_dims; i0 += 2) {
            const float ff = freq_factors ? freq_factors[i0/2] : 1.0f;
            const float th_extrap = theta / ff;
            float th = freq_scale * th_extrap, mscale = attn_factor;
            if (ext_factor != 0.0f) {
                float yv = (i0/2 - corr[0]) / fmaxf(0.001f, corr[1] - corr[0]);
                float ramp = (1.0f - fminf(1.0f, fmaxf(0.0f, yv))) * ext_factor;
                th = th * (1 - ramp) + th_extrap * ramp;
                mscale *= 1.0f + 0.1f * logf(1.0f / freq_scale);
            }
            cs[i0] = cosf(th) * mscale; cs[i0 + 1] = sinf(th) * mscale;
            theta *= theta_scale;
        }
        for (int64_t h = 0; h < n_head; ++h) {
            const float * s = x + (t*n_head + h)*head_dim; float * d = y + (t*n_head + h)*head_dim;
            for (int i0 = 0; i0 < n_dims; i0 += 2) {
                const int a = (mode & 2) ? i0/2 : i0, b = (mode & 2) ? i0/2 + n_dims/2 : i0 + 1;
                const float x0 = s[a], x1 = s[b];
                d[a] = x0*cs[i0] - x1*cs[i0 + 1];
                d[b] = x0*cs[i0 + 1] + x1*cs[i0];
            }
            for (int64_t i = n_dims; i < head_dim; ++i) d[i] = s[i];
        }
    }
    free(cs);
}

void or_swiglu(const float * gate, const float * up, float * y, int64_t n) {
    for (int64_t i = 0; i < n; ++i) y[i] = (gate[i] / (1.0f + expf(-gate[i]))) * up[i];
}

/* rows of F32 -> F16 rows scattered at idx (KV-cache write) */
void or_set_rows_f16(const float * src, const int64_t * idx, uint16_t * dst, int64_t ncols, int64_t nrows) {
    for (int64_t r = 0; r < nrows; ++r) for (int64_t i = 0; i < ncols; ++i) dst[idx[r]*ncols + i] = or_f2h(src[r*ncols + i]);
}
void or_get_rows_f32(const float * src, const int32_t * idx, float * dst, int64_t ncols, int64_t nrows) {
    for (int64_t r = 0; r < nrows; ++r) memcpy(dst + r*ncols, src + (int64_t) idx[r]*ncols, 4*ncols);
}

/* softmax over rows with optional f32 mask row (broadcast over rows by caller) */
void or_soft_max(const float * x, const float * mask, float * y, int64_t ncols, int64_t nrows, float scale) {
    for (int64_t r = 0; r < nrows; ++r, x += ncols, y += ncols) {
        float mx = -INFINITY;
        for (int64_t i = 0; i < ncols; ++i) { y[i] = x[i]*scale + (mask ? mask[r*ncols + i] : 0.0f); if (y[i] > mx) mx = y[i]; }
        double sum = 0.0;
        for (int64_t i = 0; i < ncols; ++i) { float e = expf(y[i] - mx); y[i] = e; sum += (double) e; }
        const float inv = (float) (1.0 / sum);
        for (int64_t i = 0; i < ncols; ++i) y[i] *= inv;
    }
}

/* ---- FLASH_ATTN_EXT, F16 K/V ------------------------------------------------------------------
 * q   : F32 [n_q][n_head][D]          (element (t, h, d) at q + t*q_stride_t + h*q_stride_h + d; strides in floats)
 * k,v : F16 [n_head_kv][n_kv][D]      (element (hk, c, d) at k + hk*kv_stride_h + c*kv_stride_c + d; strides in halves)
 * mask: F16 [n_q][mask_stride] or NULL (additive, -inf skips the column)
 * dst : F32 [n_q][n_head][D]
 * f16_acc = 1 reproduces the CPU backend (VKQ accumulated in F16); 0 gives the exact-f32 accumulation variant. */
void or_flash_attn_f16(const float * q, const uint16_t * k, const uint16_t * v, const uint16_t * mask, float * dst,
                       int64_t D, int64_t n_q, int64_t n_head, int64_t n_head_kv, int64_t n_kv,
                       int64_t q_stride_t, int64_t q_stride_h, int64_t kv_stride_h, int64_t kv_stride_c, int64_t mask_stride,
                       float scale, int f16_acc) {
    const int64_t gqa = n_head / n_head_kv;
    float * acc32 = malloc(sizeof(float) * D); uint16_t * acc16 = malloc(2 * D); uint16_t * q16 = malloc(2 * D);
    for (int64_t t = 0; t < n_q; ++t) for (int64_t h = 0; h < n_head; ++h) {
        const float * qr = q + t*q_stride_t + h*q_stride_h;
        for (int64_t d = 0; d < D; ++d) q16[d] = or_f2h(qr[d]);
        for (int64_t d = 0; d < D; ++d) { acc32[d] = 0.0f; acc16[d] = 0; }
        float S = 0.0f, M = -INFINITY;
        const uint16_t * kh = k + (h/gqa)*kv_stride_h, * vh = v + (h/gqa)*kv_stride_h;
        for (int64_t c = 0; c < n_kv; ++c) {
            const float mv = mask ? or_h2f(mask[t*mask_stride + c]) : 0.0f;
            if (mv == -INFINITY) continue;
            /* ggml_vec_dot_f16: f16*f16 products summed in f32 (SIMD lanes); restated with a double accumulator */
            double sd = 0.0;
            for (int64_t d = 0; d < D; ++d) sd += (double) (or_h2f(kh[c*kv_stride_c + d]) * or_h2f(q16[d]));
            float s = (float) sd * scale + mv;
            float ms = 1.0f, vs = 1.0f;
            if (s > M) { ms = expf(M - s); M = s;
                         if (f16_acc) for (int64_t d = 0; d < D; ++d) acc16[d] = or_f2h(or_h2f(acc16[d]) * ms);
                         else         for (int64_t d = 0; d < D; ++d) acc32[d] *= ms; }
            else       { vs = expf(s - M); }
            if (f16_acc) for (int64_t d = 0; d < D; ++d) acc16[d] = or_f2h(or_h2f(acc16[d]) + or_h2f(vh[c*kv_stride_c + d]) * vs);
            else         for (int64_t d = 0; d < D; ++d) acc32[d] += or_h2f(vh[c*kv_stride_c + d]) * vs;
            S = S*ms + vs;
        }
        const float inv = S == 0.0f ? 0.0f : 1.0f / S;
        float * o = dst + (t*n_head + h)*D;
        for (int64_t d = 0; d < D; ++d) o[d] = (f16_acc ? or_h2f(acc16[d]) : acc32[d]) * inv;
    }
    free(acc32); free(acc16); free(q16);
}

/* ---- ops of the APM / VPM encoder graphs ---------------------------------------------------------------------------
 * NORM     ggml/src/ggml-cpu/ops.cpp:3450-3495: mean = sum/n, y = x - mean, variance = sum(y*y)/n (ggml_vec_cvar_f32, vec.cpp:407), y *= 1/sqrtf(var+eps)
 * IM2COL   ops.cpp:6160-6301: [N, IC, IH, IW] -> [N, OH, OW, IC*KH*KW], zero padding, F32 -> F16 (or F32)
 * POOL_1D  ops.cpp:7212-7280: kernel == stride, no padding; avg = sequential f32 sum / k, max starts at -FLT_MAX */
void or_norm(const float * x, float * y, int64_t ncols, int64_t nrows, float eps) {
    for (int64_t r = 0; r < nrows; ++r, x += ncols, y += ncols) {
        double s = 0.0;
        for (int64_t i = 0; i < ncols; ++i) s += (double) x[i];
        const float mean = (float) s / ncols;
        double v = 0.0;
        for (int64_t i = 0; i < ncols; ++i) { y[i] = x[i] - mean; v += (double) (y[i] * y[i]); }
        const float variance = (float) (v / ncols);
        const float scale = 1.0f / sqrtf(variance + eps);
        for (int64_t i = 0; i < ncols; ++i) y[i] *= scale;
    }
}

void or_im2col(const float * x, void * dst, int dst_f16, int64_t N, int64_t IC, int64_t IH, int64_t IW, int64_t KH, int64_t KW, int64_t OH, int64_t OW,
               int s0, int s1, int p0, int p1, int d0, int d1) {
    for (int64_t in = 0; in < N; ++in) for (int64_t ioh = 0; ioh < OH; ++ioh) for (int64_t iow = 0; iow < OW; ++iow) for (int64_t iic = 0; iic < IC; ++iic) {
        const int64_t base = (in*OH*OW + ioh*OW + iow) * (IC*KH*KW) + iic*(KH*KW);
        const float * src = x + (in*IC + iic) * IH*IW;
        for (int64_t ikh = 0; ikh < KH; ++ikh) for (int64_t ikw = 0; ikw < KW; ++ikw) {
            const int64_t iiw = iow*s0 + ikw*d0 - p0, iih = ioh*s1 + ikh*d1 - p1;
            const float v = (iih < 0 || iih >= IH || iiw < 0 || iiw >= IW) ? 0.0f : src[iih*IW + iiw];
            if (dst_f16) ((uint16_t *) dst)[base + ikh*KW + ikw] = or_f2h(v); else ((float *) dst)[base + ikh*KW + ikw] = v;
        }
    }
}

void or_pool_1d(const float * x, float * dst, int64_t ncols_in, int64_t nrows, int op, int k) {
    const int64_t rs = ncols_in / k;
    for (int64_t r = 0; r < nrows; ++r) {
        const float * srow = x + r * ncols_in;
        float * drow = dst + r * rs;
        int64_t j = 0;
        for (int64_t i = 0; i < rs; ++i) {
            drow[i] = op == 1 ? 0.0f : -3.402823466e+38f;
            for (int ki = 0; ki < k; ++ki, ++j) { if (op == 1) drow[i] += srow[j]; else if (srow[j] > drow[i]) drow[i] = srow[j]; }
            if (op == 1) drow[i] /= k;
        }
    }
}
