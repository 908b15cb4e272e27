// flash_attn.cu — GGML_OP_FLASH_ATTN_EXT over the F16 KV cache, decode / small-batch form (split-KV, online softmax).
//
// Replaces ggml-cuda/fattn.cu:195-341 -> flash_attn_ext_vec (fattn-vec.cuh:19) + flash_attn_combine_results
// (fattn-common.cuh).  Arithmetic follows the CPU oracle (ggml-cpu/ops.cpp:7912-8148): Q rounded to f16, K.Q products of
// f16 values accumulated in f32, s = dot*scale + mask, online softmax in f32; V is accumulated in f32 here (the CPU backend
// accumulates in f16 — lossier; the reference's own backend test allows NMSE 5e-4, tests/test-backend-ops.cpp:5085).
//
// B200-first: HBM-bound on the K/V read.  One CTA owns (KV-chunk, kv-head, q-token) and serves all G query heads of the GQA
// group from ONE pass over K and V, so every cache byte is read once per token; a D-half row (256 B at D = 128) is read by
// 16 lanes x 16 B, rows whose mask is -inf are never fetched (the reference pre-scans the mask for that,
// fattn-common.cuh flash_attn_mask_to_KV_max); the grid is sized to ~2 CTAs per SM and a tiny second kernel merges the
// per-chunk (m, l, acc) partials.
#include "common.cuh"
#include <math.h>

namespace b200 {

constexpr int FA_THREADS = 256;
constexpr int FA_TILE    = 256;      // KV positions per softmax tile

struct FaArgs {
    const char * q; const char * k; const char * v; const char * mask; char * dst;
    int64_t q_nb1, q_nb2, q_nb3, k_nb1, k_nb2, k_nb3, v_nb1, v_nb2, v_nb3, m_nb1, m_nb2, m_nb3, d_nb1, d_nb2, d_nb3;
    int64_t n_q, n_kv, n_head, n_head_kv, k_ne3, m_ne2, m_ne3;
    int ratio, groups, splits; int64_t chunk;
    float scale;
    float * part_acc; float2 * part_ml;
};

struct FaPlan { int G, groups, splits; int64_t chunk; };
static FaPlan fa_plan(int64_t n_q_total, int64_t n_kv, int64_t n_head, int64_t n_head_kv) {
    FaPlan P;
    const int ratio = (int) (n_head / n_head_kv);
    P.G = ratio % 4 == 0 ? 4 : ratio % 2 == 0 ? 2 : 1;
    P.groups = ratio / P.G;
    const int64_t base = n_head_kv * P.groups * n_q_total;
    int64_t splits = (2 * (int64_t) sm_count() + base - 1) / base;
    const int64_t max_splits = (n_kv + 63) / 64;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int64_t chunk = (n_kv + splits - 1) / splits;
    chunk = (chunk + 31) / 32 * 32;
    P.chunk = chunk;
    P.splits = (int) ((n_kv + chunk - 1) / chunk);
    if (P.splits < 1) P.splits = 1;
    return P;
}

// 8 consecutive elements of a K / V row, as raw bits and as floats.  KT = B200_F16: 16 bytes of halves.  KT = B200_Q8_0 / B200_Q4_0 (a quantised KV cache, -ctk / -ctv):
// the lane's 8 codes (8 int8 bytes, or the low / high nibbles of 8 bytes) and the block's f16 d, fetched with 2-byte loads (34- / 18-byte blocks are only 2-byte aligned)
// and dequantised in registers as dequantize_row_q8_0 / q4_0 do (ggml-quants.c:390-402, 307-325) — no F16 staging pass, the cache is read once at its quantised size.
template <int KT> __device__ __forceinline__ uint4 kv_load8(const char * row, int hl) {
    if (KT == B200_F16) return ldg_stream16(row + hl * 16);
    const int e0 = hl * 8, blk = e0 >> 5, o = e0 & 31;
    const uint16_t * b = (const uint16_t *) (row + blk * (KT == B200_Q8_0 ? 34 : 18));
    const uint16_t * q = b + 1 + (KT == B200_Q8_0 ? o >> 1 : (o & 15) >> 1);
    uint4 r;
    r.x = (uint32_t) __ldg(q) | ((uint32_t) __ldg(q + 1) << 16); r.y = (uint32_t) __ldg(q + 2) | ((uint32_t) __ldg(q + 3) << 16);
    r.z = __ldg(b); r.w = 0;
    return r;
}
template <int KT> __device__ __forceinline__ void kv_floats8(const uint4 raw, int hl, float (&f)[8]) {
    if (KT == B200_F16) {
        const __half2 * h2 = (const __half2 *) &raw;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h2[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
        return;
    }
    const float d = h2f((uint16_t) raw.z);
    const uint32_t w[2] = { raw.x, raw.y };
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t byte = (w[i >> 2] >> (8 * (i & 3))) & 0xff;
        const int code = KT == B200_Q8_0 ? (int) (int8_t) byte : (int) (((hl * 8) & 16) ? byte >> 4 : byte & 15) - 8;
        f[i] = __fmul_rn((float) code, d);
    }
}

template <int D, int G, int KT>
__global__ void __launch_bounds__(FA_THREADS) k_fa_decode(const FaArgs A) {
    constexpr int LPR = D / 8;                 // lanes per K/V row (16 B each)
    constexpr int RPW = 32 / LPR;              // rows per warp-load
    constexpr int NRG = (FA_THREADS / 32) * RPW;
    constexpr int U   = 4;                     // rows in flight per lane
    static_assert(FA_TILE % (NRG * U) == 0, "tile");
    __shared__ float S[FA_TILE][G];
    __shared__ float s_corr[G], s_m[G], s_l[G];
    __shared__ float red[NRG][G * D + 4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = warp * RPW + lane / LPR, hl = lane % LPR;
    const int split = blockIdx.x;
    const int kvh = blockIdx.y / A.groups, grp = blockIdx.y % A.groups;
    const int64_t iq = blockIdx.z % A.n_q, ib = blockIdx.z / A.n_q;
    const int head0 = kvh * A.ratio + grp * G;

    float qreg[G][8];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const float * qp = (const float *) (A.q + iq * A.q_nb1 + (int64_t) (head0 + g) * A.q_nb2 + ib * A.q_nb3) + hl * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) qreg[g][i] = __half2float(__float2half_rn(qp[i]));
    }
    if (tid < G) { s_m[tid] = -INFINITY; s_l[tid] = 0.0f; }
    float acc[G][8];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[g][i] = 0.0f;

    const char * kb = A.k + (int64_t) kvh * A.k_nb2 + (ib % A.k_ne3) * A.k_nb3;
    const char * vb = A.v + (int64_t) kvh * A.v_nb2 + (ib % A.k_ne3) * A.v_nb3;
    const __half * mrow = A.mask ? (const __half *) (A.mask + iq * A.m_nb1 + ((int64_t) head0 % A.m_ne2) * A.m_nb2 + (ib % A.m_ne3) * A.m_nb3) : nullptr;
    const int64_t c0 = (int64_t) split * A.chunk, c1 = min(c0 + A.chunk, A.n_kv);
    __syncthreads();

    for (int64_t t0 = c0; t0 < c1; t0 += FA_TILE) {
        // ---- S = scale * K.q + mask for the tile --------------------------------------------------------------------
        for (int j0 = rg; j0 < FA_TILE; j0 += NRG * U) {
            uint4 kk[U]; float mv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t pos = t0 + j0 + u * NRG;
                mv[u] = pos < c1 ? (mrow ? __half2float(mrow[pos]) : 0.0f) : -INFINITY;
                kk[u] = make_uint4(0, 0, 0, 0);
                if (mv[u] != -INFINITY) kk[u] = kv_load8<KT>(kb + pos * A.k_nb1, hl);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float kf[8];
                kv_floats8<KT>(kk[u], hl, kf);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    float d = 0.0f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) d = fmaf(kf[i], qreg[g][i], d);
#pragma unroll
                    for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                    if (hl == 0) S[j0 + u * NRG][g] = mv[u] != -INFINITY ? d * A.scale + mv[u] : -INFINITY;
                }
            }
        }
        __syncthreads();
        // ---- online softmax bookkeeping: warp g owns head g ------------------------------------------------------------
        if (warp < G) {
            const int g = warp;
            float mx = -INFINITY;
            for (int j = lane; j < FA_TILE; j += 32) mx = fmaxf(mx, S[j][g]);
            mx = warp_max(mx);
            const float m_old = s_m[g], m_new = fmaxf(m_old, mx);
            float sum = 0.0f;
            for (int j = lane; j < FA_TILE; j += 32) {
                const float s = S[j][g];
                const float p = s == -INFINITY ? 0.0f : expf(s - m_new);
                S[j][g] = p; sum += p;
            }
            sum = warp_sum(sum);
            if (lane == 0) {
                const float corr = m_old == -INFINITY ? 1.0f : expf(m_old - m_new);     // m_new >= m_old; acc is 0 while m_old = -inf
                s_corr[g] = corr; s_m[g] = m_new; s_l[g] = s_l[g] * corr + sum;
            }
        }
        __syncthreads();
        // ---- acc = acc*corr + P.V ------------------------------------------------------------------------------------
#pragma unroll
        for (int g = 0; g < G; ++g) { const float c = s_corr[g];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[g][i] *= c; }
        for (int j0 = rg; j0 < FA_TILE; j0 += NRG * U) {
            uint4 vv[U]; float pv[U][G];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * NRG;
                bool any = false;
#pragma unroll
                for (int g = 0; g < G; ++g) { pv[u][g] = S[j][g]; any |= pv[u][g] != 0.0f; }
                vv[u] = make_uint4(0, 0, 0, 0);
                if (any) vv[u] = kv_load8<KT>(vb + (t0 + j) * A.v_nb1, hl);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float vf[8];
                kv_floats8<KT>(vv[u], hl, vf);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int g = 0; g < G; ++g) acc[g][i] = fmaf(pv[u][g], vf[i], acc[g][i]);
            }
        }
        __syncthreads();
    }
    // ---- reduce the NRG row-group accumulators ------------------------------------------------------------------------
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int i = 0; i < 8; ++i) red[rg][g * D + hl * 8 + i] = acc[g][i];
    __syncthreads();
    const int64_t z = blockIdx.z;
    for (int o = tid; o < G * D; o += FA_THREADS) {
        float v = 0.0f;
#pragma unroll
        for (int r = 0; r < NRG; ++r) v += red[r][o];
        const int g = o / D, d = o % D;
        const int head = head0 + g;
        if (A.splits == 1) {
            const float l = s_l[g];
            *(float *) (A.dst + (int64_t) d * 4 + (int64_t) head * A.d_nb1 + iq * A.d_nb2 + ib * A.d_nb3) = l == 0.0f ? 0.0f : v / l;
        } else {
            const int64_t slot = (z * A.n_head + head) * A.splits + split;
            A.part_acc[slot * D + d] = v;
            if (d == 0) A.part_ml[slot] = make_float2(s_m[g], s_l[g]);
        }
    }
}

template <int D>
__global__ void __launch_bounds__(D) k_fa_combine(const FaArgs A) {
    const int head = blockIdx.x; const int64_t z = blockIdx.y;
    const int64_t iq = z % A.n_q, ib = z / A.n_q;
    const int64_t slot0 = (z * A.n_head + head) * A.splits;
    float M = -INFINITY;
    for (int s = 0; s < A.splits; ++s) M = fmaxf(M, A.part_ml[slot0 + s].x);
    float L = 0.0f, v = 0.0f;
    for (int s = 0; s < A.splits; ++s) {
        const float2 ml = A.part_ml[slot0 + s];
        if (ml.x == -INFINITY) continue;
        const float w = expf(ml.x - M);
        L += ml.y * w;
        v += A.part_acc[(slot0 + s) * D + threadIdx.x] * w;
    }
    *(float *) (A.dst + (int64_t) threadIdx.x * 4 + (int64_t) head * A.d_nb1 + iq * A.d_nb2 + ib * A.d_nb3) = L == 0.0f ? 0.0f : v / L;
}

template <int D, int KT>
static int fa_launch(const FaArgs & A, int G, int64_t nz, cudaStream_t st) {
    dim3 grid((unsigned) A.splits, (unsigned) (A.n_head_kv * A.groups), (unsigned) nz);
    if      (G == 4) k_fa_decode<D, 4, KT><<<grid, FA_THREADS, 0, st>>>(A);
    else if (G == 2) k_fa_decode<D, 2, KT><<<grid, FA_THREADS, 0, st>>>(A);
    else             k_fa_decode<D, 1, KT><<<grid, FA_THREADS, 0, st>>>(A);
    B200_LAUNCH_CHECK();
    if (A.splits > 1) {
        k_fa_combine<D><<<dim3((unsigned) A.n_head, (unsigned) nz), D, 0, st>>>(A);
        B200_LAUNCH_CHECK();
    }
    return B200_OK;
}

int dequant_rows(const b200_tensor * src, const b200_tensor * idx, const b200_tensor * dst, cudaStream_t st);     // quant_rows.cu
// fa_prefill.cu: tensor-core tiles for >= 16 query tokens
bool   fa_prefill_supported(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask, const b200_tensor * dst);
size_t fa_prefill_scratch_bytes(const b200_tensor * q, const b200_tensor * mask_or_null, int64_t m_ne3);
int    fa_prefill(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask, const b200_tensor * dst, float scale,
                  void * scratch, cudaStream_t st, void * tiles);

} // namespace b200

using namespace b200;

extern "C" int b200_flash_attn_supported(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask,
                                         const b200_tensor * dst) {
    if (!q || !k || !v || !dst) return 0;
    // quantised K / V (-ctk / -ctv q8_0 | q4_0) are staged as F16 in the scratch first, like launch_fattn's to_fp16 pass (fattn-common.cuh); a quantised V needs a
    // quantised K, because b200_flash_attn_scratch_bytes sizes the staging from K alone
    const bool kq = k->type == B200_Q8_0 || k->type == B200_Q4_0, vq = v->type == B200_Q8_0 || v->type == B200_Q4_0;
    if (q->type != B200_F32 || (k->type != B200_F16 && !kq) || (v->type != B200_F16 && !vq) || (vq && !kq) || dst->type != B200_F32) return 0;
    const int64_t D = q->ne[0];
    if ((D != 64 && D != 128) || k->ne[0] != D || v->ne[0] != D) return 0;
    if (k->ne[1] != v->ne[1] || k->ne[2] != v->ne[2] || k->ne[3] != v->ne[3] || k->ne[2] == 0 || k->ne[3] == 0) return 0;
    if (q->ne[2] % k->ne[2] != 0 || q->ne[3] % k->ne[3] != 0) return 0;
    if (q->nb[0] != 4 || k->nb[0] != type_size(k->type) || v->nb[0] != type_size(v->type) || dst->nb[0] != 4) return 0;
    if (!kq && ((uintptr_t) k->data % 16 || (k->nb[1] | k->nb[2] | k->nb[3]) % 16)) return 0;
    if (!vq && ((uintptr_t) v->data % 16 || (v->nb[1] | v->nb[2] | v->nb[3]) % 16)) return 0;
    if (kq && (((uintptr_t) k->data | k->nb[1] | k->nb[2] | k->nb[3]) & 1)) return 0;
    if (vq && (((uintptr_t) v->data | v->nb[1] | v->nb[2] | v->nb[3]) & 1)) return 0;
    if (dst->ne[0] != D || dst->ne[1] != q->ne[2] || dst->ne[2] != q->ne[1] || dst->ne[3] != q->ne[3]) return 0;
    if (mask) {
        if (mask->type != B200_F16 || mask->nb[0] != 2 || mask->ne[0] != k->ne[1] || mask->ne[1] < q->ne[1]) return 0;
        if (mask->ne[2] == 0 || mask->ne[3] == 0 || q->ne[2] % mask->ne[2] || q->ne[3] % mask->ne[3]) return 0;
        if (mask->ne[2] != 1) return 0;            // per-head masks (ALiBi-style) are not on this path
    }
    return 1;
}

// split-KV decode kernel: part_acc + part_ml of every (query, head, split)
static size_t fa_decode_scratch_bytes(const b200_tensor * q, const b200_tensor * k) {
    const int64_t nz = q->ne[1] * q->ne[3];
    const FaPlan P = fa_plan(nz, k->ne[1], q->ne[2], k->ne[2]);
    if (P.splits <= 1) return 0;
    const size_t slots = (size_t) nz * q->ne[2] * P.splits;
    return slots * q->ne[0] * 4 + slots * 8 + 16;
}

// Which kernel runs is decided in b200_flash_attn from alignment as well (fa_prefill_supported), which this function cannot see: it
// returns the LARGER of the two requirements, so the buffer fits whichever path is taken.
static size_t fa_stage_bytes(const b200_tensor * k) {                       // one F16 copy of a quantised K (and the same again for V)
    return ((size_t) (k->ne[0] * k->ne[1] * k->ne[2] * k->ne[3]) * 2 + 255) & ~(size_t) 255;
}
extern "C" size_t b200_flash_attn_scratch_bytes(const b200_tensor * q, const b200_tensor * k) {
    const int64_t nz = q->ne[1] * q->ne[3];
    if (nz == 0 || k->ne[1] == 0 || k->ne[2] == 0) return 0;
    const size_t stage = k->type != B200_F16 ? 2 * fa_stage_bytes(k) : 0;
    const size_t dec = nz <= 65535 ? fa_decode_scratch_bytes(q, k) : 0;
    if (q->ne[1] >= 16) { const size_t pre = fa_prefill_scratch_bytes(q, nullptr, q->ne[3]); return stage + (pre > dec ? pre : dec); }   // upper bound for the KV-tile counts (mask batch <= q batch)
    return stage + dec;
}

static int flash_attn_impl(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask, const b200_tensor * dst, float scale, float max_bias,
                           float logit_softcap, void * scratch, size_t scratch_bytes, void * tiles, void * stream);
extern "C" int b200_flash_attn(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask,
                               const b200_tensor * dst, float scale, float max_bias, float logit_softcap, void * scratch,
                               size_t scratch_bytes, void * stream) {
    return flash_attn_impl(q, k, v, mask, dst, scale, max_bias, logit_softcap, scratch, scratch_bytes, nullptr, stream);
}
// The [n_q, n_head * 128] result handed to the MUL_MAT that follows (wo) as prepared F16 activation tiles in `tiles` (that MUL_MAT's scratch; call it with
// B200_MM_REUSE_ACT) instead of F32 in dst->data (which is not written).  Only the tcgen05 prefill kernel does this: B200_ERR_UNSUPPORTED otherwise, and the
// caller falls back to b200_flash_attn + the MUL_MAT's own conversion pass.  `scratch` (mask pre-scan) must not overlap `tiles`.
extern "C" int b200_flash_attn_tiles(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask,
                                     const b200_tensor * dst, float scale, void * scratch, size_t scratch_bytes, void * tiles, void * stream) {
    if (!tiles || (uintptr_t) tiles % 16 || q->ne[3] != 1) return B200_ERR_UNSUPPORTED;
    return flash_attn_impl(q, k, v, mask, dst, scale, 0.0f, 0.0f, scratch, scratch_bytes, tiles, stream);
}
static int flash_attn_impl(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask, const b200_tensor * dst, float scale, float max_bias,
                           float logit_softcap, void * scratch, size_t scratch_bytes, void * tiles, void * stream) {
    if (!b200_flash_attn_supported(q, k, v, mask, dst) || max_bias != 0.0f || logit_softcap != 0.0f) return B200_ERR_UNSUPPORTED;
    const int64_t nz = q->ne[1] * q->ne[3];
    if (nz == 0 || q->ne[2] == 0) return B200_OK;
    b200_tensor kf, vf;
    // decode class (< 16 query tokens) with K and V of the SAME quantised type: k_fa_decode dequantises in its loads, nothing is staged
    const bool direct_q = k->type != B200_F16 && k->type == v->type && q->ne[1] < 16 && nz <= 65535;
    if (k->type != B200_F16 && k->ne[1] > 0 && !direct_q) {                   // stage the quantised cache views as contiguous F16 [D, n_kv, n_head_kv, n_b]
        const size_t sb = fa_stage_bytes(k);
        if (!scratch || (uintptr_t) scratch % 16 || scratch_bytes < 2 * sb) return B200_ERR_ARG;
        auto stage = [&](const b200_tensor * t, b200_tensor & f, char * where) -> int {
            if (t->type == B200_F16) { f = *t; return B200_OK; }
            f = *t; f.data = where; f.type = B200_F16; f.layout = B200_LAYOUT_NATIVE;
            f.nb[0] = 2; f.nb[1] = t->ne[0] * 2; f.nb[2] = f.nb[1] * t->ne[1]; f.nb[3] = f.nb[2] * t->ne[2];
            return dequant_rows(t, nullptr, &f, (cudaStream_t) stream);
        };
        int rc = stage(k, kf, (char *) scratch); if (rc) return rc;
        rc = stage(v, vf, (char *) scratch + sb); if (rc) return rc;
        k = &kf; v = &vf;
        scratch = (char *) scratch + 2 * sb; scratch_bytes -= 2 * sb;
    }
    if (k->ne[1] > 0 && fa_prefill_supported(q, k, v, mask, dst)) {
        const bool have = scratch && scratch_bytes >= fa_prefill_scratch_bytes(q, mask, mask ? mask->ne[3] : 1) && (uintptr_t) scratch % 4 == 0;
        return fa_prefill(q, k, v, mask, dst, scale, have ? scratch : nullptr, (cudaStream_t) stream, tiles);
    }
    if (tiles) return B200_ERR_UNSUPPORTED;
    if (nz > 65535) return B200_ERR_UNSUPPORTED;
    const FaPlan P = fa_plan(nz, k->ne[1], q->ne[2], k->ne[2]);
    FaArgs A = {};
    A.q = (const char *) q->data; A.k = (const char *) k->data; A.v = (const char *) v->data; A.mask = mask ? (const char *) mask->data : nullptr;
    A.dst = (char *) dst->data;
    A.q_nb1 = q->nb[1]; A.q_nb2 = q->nb[2]; A.q_nb3 = q->nb[3];
    A.k_nb1 = k->nb[1]; A.k_nb2 = k->nb[2]; A.k_nb3 = k->nb[3];
    A.v_nb1 = v->nb[1]; A.v_nb2 = v->nb[2]; A.v_nb3 = v->nb[3];
    if (mask) { A.m_nb1 = mask->nb[1]; A.m_nb2 = mask->nb[2]; A.m_nb3 = mask->nb[3]; A.m_ne2 = mask->ne[2]; A.m_ne3 = mask->ne[3]; } else { A.m_ne2 = A.m_ne3 = 1; }
    A.d_nb1 = dst->nb[1]; A.d_nb2 = dst->nb[2]; A.d_nb3 = dst->nb[3];
    A.n_q = q->ne[1]; A.n_kv = k->ne[1]; A.n_head = q->ne[2]; A.n_head_kv = k->ne[2]; A.k_ne3 = k->ne[3];
    A.ratio = (int) (q->ne[2] / k->ne[2]); A.groups = P.groups; A.splits = P.splits; A.chunk = P.chunk; A.scale = scale;
    if (k->ne[1] == 0) { A.splits = 1; A.chunk = 32; }
    if (A.splits > 1) {
        if (!scratch || scratch_bytes < fa_decode_scratch_bytes(q, k) || (uintptr_t) scratch % 16) return B200_ERR_ARG;
        const size_t slots = (size_t) nz * q->ne[2] * P.splits;
        A.part_acc = (float *) scratch;
        A.part_ml  = (float2 *) ((char *) scratch + ((slots * q->ne[0] * 4 + 15) & ~(size_t) 15));
    }
    cudaStream_t st = (cudaStream_t) stream;
    if (k->type == B200_Q8_0) return q->ne[0] == 128 ? fa_launch<128, B200_Q8_0>(A, P.G, nz, st) : fa_launch<64, B200_Q8_0>(A, P.G, nz, st);
    if (k->type == B200_Q4_0) return q->ne[0] == 128 ? fa_launch<128, B200_Q4_0>(A, P.G, nz, st) : fa_launch<64, B200_Q4_0>(A, P.G, nz, st);
    return q->ne[0] == 128 ? fa_launch<128, B200_F16>(A, P.G, nz, st) : fa_launch<64, B200_F16>(A, P.G, nz, st);
}
