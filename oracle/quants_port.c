/* oracle/quants_port.c — TEST INFRASTRUCTURE (CPU oracle), never linked into the product.
 *
 * Plain-C restatement of the reference's quantised MUL_MAT arithmetic (the decode hot path):
 *   - weight block decode           ggml/src/ggml-quants.c  q4_0 :307-325  q8_0 :390-402  q4_K :1352-1374
 *                                   q5_K :1554-1578  q6_K :1762-1791  scale unpack :703-711
 *   - activation quantisation       quantize_row_q8_0_ref ggml-quants.c:199-222 (+ AVX2 variant
 *                                   ggml-cpu/arch/x86/quants.c:297-360), quantize_row_q8_K_ref :2555-2592
 *   - integer dot products          ggml-cpu/quants.c  q4_0 :115-149  q8_0 :305-333  q4_K :550-623
 *                                   q5_K :625-703  q6_K :705-758
 *   - MUL_MAT driver                ggml-cpu/ggml-cpu.c:1210-1402 (src1 rows -> vec_dot_type, one vec_dot per (row, col))
 *
 * Pinned against the reference itself (oracle/_ref: libggml-base.so / libggml-cpu.so) by
 * tests/test_oracle_pin.py and against tests/golden/*.npz.
 *
 * Float accumulation order: per (super-)block, left to right.  The reference's generic C code keeps eight
 * interleaved float lanes and its AVX2 code yet another order; the INTEGER parts (sub-block dot products,
 * bsum*min terms) are identical in all of them, so the variants agree to a few ulp of the running sum.
 */
#include "blocks.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

/* ---- 6-bit (scale, min) pairs of q4_K / q5_K : 12 bytes -> 8 scales + 8 mins ------------------ */
static void kquant_scale_min(const uint8_t * p, int j, int * sc, int * mn) {
    if (j < 4) { *sc = p[j] & 63;  *mn = p[j + 4] & 63; }
    else       { *sc = (p[j + 4] & 15) | ((p[j - 4] >> 6) << 4);
                 *mn = (p[j + 4] >> 4) | ((p[j]     >> 6) << 4); }
}

/* ---- integer view of one weight: returns the unsigned/signed code of element e of a block ------ */
static inline int q4_0_code(const or_q4_0 * b, int e) { return (e < 16 ? (b->qs[e] & 15) : (b->qs[e - 16] >> 4)) - 8; }
static inline int q4_K_code(const or_q4_K * b, int e) { int g = e >> 6, l = e & 31; return (e & 32) ? b->qs[32*g + l] >> 4 : b->qs[32*g + l] & 15; }
static inline int q5_K_code(const or_q5_K * b, int e) {
    int g = e >> 6, l = e & 31, hi = (b->qh[l] >> (e >> 5)) & 1;
    return ((e & 32) ? b->qs[32*g + l] >> 4 : b->qs[32*g + l] & 15) + 16*hi;
}
static inline int q6_K_code(const or_q6_K * b, int e) {
    int h = e >> 7, r = e & 127, l = r & 31, quad = r >> 5;               /* quad: 0..3 -> (ql lo/hi nibble, qh shift) */
    int lo = (quad & 1) ? b->ql[64*h + 32 + l] : b->ql[64*h + l];
    lo = (quad & 2) ? lo >> 4 : lo & 15;
    return (lo | (((b->qh[32*h + l] >> (2*quad)) & 3) << 4)) - 32;
}

/* ---- dequantise rows ------------------------------------------------------------------------- */
void or_dequant_q4_0(const void * vx, float * y, int64_t n) {
    const or_q4_0 * x = vx;
    for (int64_t i = 0; i < n; ++i) y[i] = q4_0_code(&x[i/32], i%32) * or_h2f(x[i/32].d);
}
void or_dequant_q8_0(const void * vx, float * y, int64_t n) {
    const or_q8_0 * x = vx;
    for (int64_t i = 0; i < n; ++i) y[i] = x[i/32].qs[i%32] * or_h2f(x[i/32].d);
}
void or_dequant_q4_K(const void * vx, float * y, int64_t n) {
    const or_q4_K * x = vx;
    for (int64_t i = 0; i < n; ++i) {
        const or_q4_K * b = &x[i/256]; int e = i%256, sc, mn;
        kquant_scale_min(b->sc, e/32, &sc, &mn);
        const float d = or_h2f(b->d) * sc, m = or_h2f(b->dmin) * mn;
        y[i] = d * q4_K_code(b, e) - m;
    }
}
void or_dequant_q5_K(const void * vx, float * y, int64_t n) {
    const or_q5_K * x = vx;
    for (int64_t i = 0; i < n; ++i) {
        const or_q5_K * b = &x[i/256]; int e = i%256, sc, mn;
        kquant_scale_min(b->sc, e/32, &sc, &mn);
        const float d = or_h2f(b->d) * sc, m = or_h2f(b->dmin) * mn;
        y[i] = d * q5_K_code(b, e) - m;
    }
}
void or_dequant_q6_K(const void * vx, float * y, int64_t n) {
    const or_q6_K * x = vx;
    for (int64_t i = 0; i < n; ++i) {
        const or_q6_K * b = &x[i/256]; int e = i%256;
        y[i] = or_h2f(b->d) * b->sc[e/16] * (int8_t) q6_K_code(b, e);
    }
}

/* ---- activation quantisers ------------------------------------------------------------------- */
/* variant 0: scalar reference (roundf, id = 1/d); variant 1: the x86 SIMD build (nearest-even, id = 127/amax) */
void or_quantize_q8_0(const float * x, void * vy, int64_t n, int variant) {
    or_q8_0 * y = vy;
    for (int64_t b = 0; b < n/32; ++b) {
        float amax = 0.0f;
        for (int j = 0; j < 32; ++j) { float a = fabsf(x[32*b + j]); if (a > amax) amax = a; }
        const float d  = amax / 127.0f;
        const float id = variant ? (amax != 0.0f ? 127.0f / amax : 0.0f) : (d != 0.0f ? 1.0f / d : 0.0f);
        y[b].d = or_f2h(d);
        for (int j = 0; j < 32; ++j) {
            const float v = x[32*b + j] * id;
            y[b].qs[j] = (int8_t) (variant ? nearbyintf(v) : roundf(v));
        }
    }
}
/* quantize_row_q4_0_ref, ggml-quants.c:36-71 (what SET_ROWS / CPY into a q4_0 KV cache run on the CPU: type_traits_cpu.from_float = quantize_row_q4_0 -> _ref) */
void or_quantize_q4_0(const float * x, void * vy, int64_t n) {
    or_q4_0 * y = vy;
    for (int64_t b = 0; b < n/32; ++b) {
        float amax = 0.0f, max = 0.0f;                        /* the FIRST element of largest magnitude keeps its sign */
        for (int j = 0; j < 32; ++j) { const float v = x[32*b + j]; if (amax < fabsf(v)) { amax = fabsf(v); max = v; } }
        const float d  = max / -8;
        const float id = d ? 1.0f/d : 0.0f;
        y[b].d = or_f2h(d);
        for (int j = 0; j < 16; ++j) {
            const float x0 = x[32*b + j]*id, x1 = x[32*b + 16 + j]*id;
            int q0 = (int8_t) (x0 + 8.5f), q1 = (int8_t) (x1 + 8.5f);
            if (q0 > 15) q0 = 15;
            if (q1 > 15) q1 = 15;
            y[b].qs[j] = (uint8_t) (q0 | (q1 << 4));
        }
    }
}
void or_quantize_q8_K(const float * x, void * vy, int64_t n) {
    or_q8_K * y = vy;
    for (int64_t b = 0; b < n/256; ++b, x += 256) {
        float amax = 0.0f, vmax = 0.0f;                       /* vmax keeps the SIGN of the largest-magnitude element */
        for (int j = 0; j < 256; ++j) { float a = fabsf(x[j]); if (a > amax) { amax = a; vmax = x[j]; } }
        if (amax == 0.0f) { memset(&y[b], 0, sizeof(or_q8_K)); continue; }
        const float iscale = -127.0f / vmax;
        for (int j = 0; j < 256; ++j) {
            int v = (int) nearbyintf(iscale * x[j]);          /* reference uses the 12582912.f magic add == nearest-even */
            y[b].qs[j] = (int8_t) (v > 127 ? 127 : v);
        }
        for (int g = 0; g < 16; ++g) { int s = 0; for (int j = 0; j < 16; ++j) s += y[b].qs[16*g + j]; y[b].bsums[g] = (int16_t) s; }
        y[b].d = 1.0f / iscale;
    }
}

/* ---- dot products: weights row (n elements) x pre-quantised activation row ---------------------- */
float or_vec_dot_q4_0_q8_0(int64_t n, const void * vw, const void * va) {
    const or_q4_0 * w = vw; const or_q8_0 * a = va; float acc = 0.0f;
    for (int64_t b = 0; b < n/32; ++b) {
        int s = 0; for (int e = 0; e < 32; ++e) s += q4_0_code(&w[b], e) * a[b].qs[e];
        acc += s * or_h2f(w[b].d) * or_h2f(a[b].d);
    }
    return acc;
}
float or_vec_dot_q8_0_q8_0(int64_t n, const void * vw, const void * va) {
    const or_q8_0 * w = vw; const or_q8_0 * a = va; float acc = 0.0f;
    for (int64_t b = 0; b < n/32; ++b) {
        int s = 0; for (int e = 0; e < 32; ++e) s += w[b].qs[e] * a[b].qs[e];
        acc += s * (or_h2f(w[b].d) * or_h2f(a[b].d));
    }
    return acc;
}
/* q4_K and q5_K share the algebra:  sum_j sc_j * <code, q8>_j  and  sum_j mn_j * bsum_j  */
static float kquant45_block(const uint8_t * scales, uint16_t hd, uint16_t hdmin, const or_q8_K * a,
                            const void * blk, int is5) {
    int isum = 0, msum = 0;
    for (int j = 0; j < 8; ++j) {
        int sc, mn, dot = 0; kquant_scale_min(scales, j, &sc, &mn);
        for (int e = 32*j; e < 32*j + 32; ++e)
            dot += (is5 ? q5_K_code(blk, e) : q4_K_code(blk, e)) * a->qs[e];
        isum += sc * dot;
        msum += mn * (a->bsums[2*j] + a->bsums[2*j + 1]);
    }
    return (or_h2f(hd) * a->d) * isum - (or_h2f(hdmin) * a->d) * msum;
}
float or_vec_dot_q4_K_q8_K(int64_t n, const void * vw, const void * va) {
    const or_q4_K * w = vw; const or_q8_K * a = va; float acc = 0.0f;
    for (int64_t b = 0; b < n/256; ++b) acc += kquant45_block(w[b].sc, w[b].d, w[b].dmin, &a[b], &w[b], 0);
    return acc;
}
float or_vec_dot_q5_K_q8_K(int64_t n, const void * vw, const void * va) {
    const or_q5_K * w = vw; const or_q8_K * a = va; float acc = 0.0f;
    for (int64_t b = 0; b < n/256; ++b) acc += kquant45_block(w[b].sc, w[b].d, w[b].dmin, &a[b], &w[b], 1);
    return acc;
}
float or_vec_dot_q6_K_q8_K(int64_t n, const void * vw, const void * va) {
    const or_q6_K * w = vw; const or_q8_K * a = va; float acc = 0.0f;
    for (int64_t b = 0; b < n/256; ++b) {
        int isum = 0;
        for (int g = 0; g < 16; ++g) {
            int dot = 0; for (int e = 16*g; e < 16*g + 16; ++e) dot += q6_K_code(&w[b], e) * a[b].qs[e];
            isum += w[b].sc[g] * dot;
        }
        acc += (or_h2f(w[b].d) * a[b].d) * isum;
    }
    return acc;
}

/* ---- type tables ----------------------------------------------------------------------------- */
size_t or_row_size(int type, int64_t k) {
    switch (type) {
        case OR_F32:  return 4*k;            case OR_F16:  return 2*k;          case OR_BF16: return 2*k;
        case OR_Q4_0: return k/32*18;        case OR_Q8_0: return k/32*34;
        case OR_Q4_K: return k/256*144;      case OR_Q5_K: return k/256*176;    case OR_Q6_K: return k/256*210;
        case OR_Q8_K: return k/256*292;
    }
    return 0;
}
int or_dequant_row(int type, const void * x, float * y, int64_t n) {
    switch (type) {
        case OR_Q4_0: or_dequant_q4_0(x, y, n); return 0;   case OR_Q8_0: or_dequant_q8_0(x, y, n); return 0;
        case OR_Q4_K: or_dequant_q4_K(x, y, n); return 0;   case OR_Q5_K: or_dequant_q5_K(x, y, n); return 0;
        case OR_Q6_K: or_dequant_q6_K(x, y, n); return 0;
        case OR_F32:  memcpy(y, x, 4*n); return 0;
        case OR_F16:  for (int64_t i = 0; i < n; ++i) y[i] = or_h2f(((const uint16_t *) x)[i]); return 0;
        case OR_BF16: for (int64_t i = 0; i < n; ++i) { uint32_t u = (uint32_t)((const uint16_t *) x)[i] << 16; memcpy(&y[i], &u, 4); } return 0;
    }
    return -1;
}

/* ---- MUL_MAT  dst[m, n] = W[m, k] . X[n, k]^T  (ggml: a = [k, m] weights, b = [k, n] F32, dst = [m, n] F32) --------
 * Activations are first converted to the weight type's vec_dot_type exactly like the CPU backend does
 * (ggml-cpu.c:196-350 table; :1281-1330 conversion): q4_0/q8_0 -> q8_0, K-quants -> q8_K, f16 -> f16, f32 -> f32.
 * q8_0_variant selects the scalar (0) or SIMD (1) q8_0 rounding flavour.  Returns 0 on success. */
int or_mul_mat(int type, const void * W, const float * X, float * dst, int64_t m, int64_t k, int64_t n, int q8_0_variant) {
    const size_t wrow = or_row_size(type, k);
    if (!wrow) return -1;
    for (int64_t c = 0; c < n; ++c) {
        const float * x = X + c*k;
        void * act = NULL;
        if (type == OR_Q4_0 || type == OR_Q8_0) { act = malloc(or_row_size(OR_Q8_0, k)); or_quantize_q8_0(x, act, k, q8_0_variant); }
        else if (type == OR_Q4_K || type == OR_Q5_K || type == OR_Q6_K) { act = malloc(or_row_size(OR_Q8_K, k)); or_quantize_q8_K(x, act, k); }
        for (int64_t r = 0; r < m; ++r) {
            const char * w = (const char *) W + r*wrow;
            float v = 0.0f;
            switch (type) {
                case OR_Q4_0: v = or_vec_dot_q4_0_q8_0(k, w, act); break;
                case OR_Q8_0: v = or_vec_dot_q8_0_q8_0(k, w, act); break;
                case OR_Q4_K: v = or_vec_dot_q4_K_q8_K(k, w, act); break;
                case OR_Q5_K: v = or_vec_dot_q5_K_q8_K(k, w, act); break;
                case OR_Q6_K: v = or_vec_dot_q6_K_q8_K(k, w, act); break;
                case OR_F32:  { double s = 0; for (int64_t i = 0; i < k; ++i) s += (double)((const float *) w)[i] * x[i]; v = (float) s; } break;
                case OR_F16:  { /* CPU rounds the activation to f16 first (vec_dot_type F16), accumulates in f32 */
                                double s = 0; for (int64_t i = 0; i < k; ++i) s += (double) or_h2f(((const uint16_t *) w)[i]) * or_h2f(or_f2h(x[i])); v = (float) s; } break;
                case OR_BF16: { double s = 0; for (int64_t i = 0; i < k; ++i) {
                                    uint32_t u = (uint32_t)((const uint16_t *) w)[i] << 16; float wf; memcpy(&wf, &u, 4);
                                    uint32_t xb; memcpy(&xb, &x[i], 4); xb = (xb + (0x7fff + ((xb >> 16) & 1))) & 0xffff0000u; float xf; memcpy(&xf, &xb, 4);
                                    s += (double) wf * xf; } v = (float) s; } break;
                default: free(act); return -1;
            }
            dst[c*m + r] = v;
        }
        free(act);
    }
    return 0;
}
