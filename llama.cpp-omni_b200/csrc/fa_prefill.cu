// fa_prefill.cu — GGML_OP_FLASH_ATTN_EXT for many query tokens (prefill / batched decode): tiled online-softmax attention on tensor cores.
//
// Replaces flash_attn_ext_f16 (ggml-cuda/fattn-mma-f16.cuh:1246, dispatch fattn.cu:195-341) and the reference's mask pre-scan
// flash_attn_mask_to_KV_max (fattn-common.cuh).  Oracle: ggml-cpu/ops.cpp:7912-8148 (Q rounded to f16, f32 softmax; the CPU accumulates V in
// f16, we accumulate in f32 — the reference's own bar for this op is NMSE <= 5e-4, tests/test-backend-ops.cpp:5085).
//
// One CTA = 128 query tokens of one head (8 warps x 16 rows); K/V tiles of 64 positions stream through a double-buffered, XOR-swizzled
// shared-memory ring with cp.async (16 B per thread, rows of the F16 cache are 256 B = 16 chunks); S = Q.K^T and O += P.V run as
// mma.sync.m16n8k16 f16 -> f32 with ldmatrix(.trans) operand fetch, the online softmax lives in the accumulator registers (FlashAttention-2
// layout: S accumulators of two 8-wide n-blocks ARE the A fragment of the P.V product).  Attention is ~4 % of the prefill FLOPs (SURVEY §8d:
// 1.24 of 29.7 TFLOP at 2048 tokens), the tcgen05 budget of round 1 went to the 96 % in k_mmq_tc; a TMEM-resident FA is the next step.
//   * k_fa_kvmax: per query tile (FP_BM rows), the number of KV tiles that contain any unmasked position (causal masks: everything right of the
//     diagonal is skipped, which halves the work) — the mask is shared by all heads, so this runs once per launch, not per head.
#include "common.cuh"
#include <math.h>

namespace b200 {

constexpr int FP_BM = 128, FP_BN = 64, FP_THREADS = 256;     // 8 warps x 16 query rows share every K/V tile

struct FaPArgs {
    const char * q; const char * k; const char * v; const char * mask; char * dst;
    int64_t q_nb1, q_nb2, q_nb3, k_nb1, k_nb2, k_nb3, v_nb1, v_nb2, v_nb3, m_nb1, m_nb3, d_nb1, d_nb2, d_nb3;
    int64_t n_q, n_kv, n_head, n_head_kv, k_ne3, m_ne3;
    float scale;
    const int32_t * kv_tiles;      // [n_batch_mask][n_q_tiles] KV tiles to visit, or null = all
    const int32_t * kv_plain;      // [n_batch_mask][n_q_tiles] leading KV tiles whose mask is all 0 for the tile's rows (no mask loads needed), or null
};

__device__ __forceinline__ uint32_t fp_smem(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void fp_cp16(uint32_t dst, const void * src, bool valid) {
    const int sz = valid ? 16 : 0;                                   // src-size 0: the 16 bytes are zero-filled (rows past n_kv)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void fp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void fp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fp_ldsm4(uint32_t a, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void fp_ldsm4t(uint32_t a, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void fp_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t fp_pack(float lo, float hi) { const __half2 h = __floats2half2_rn(lo, hi); return *(const uint32_t *) &h; }

// per query tile: (1) how many KV tiles must be visited = 1 + the last tile that holds an unmasked (> -inf) position for any of its rows;
// (2) how many LEADING tiles have an all-zero mask for every row (for a causal mask: everything left of the diagonal tile) — those need no
// mask loads at all.  One CTA per (query tile, KV tile) pair, 16-byte mask loads, atomicMax / atomicMin into out[] (initialised by the
// launcher): the 8 MB mask of a 2048-token prompt is read once, in parallel.
__global__ void __launch_bounds__(128) k_fa_kvmax(const char * mask, int64_t m_nb1, int64_t m_nb3, int64_t n_q, int64_t n_kv, int32_t * out_max, int32_t * out_plain, int n_q_tiles) {
    const int qt = blockIdx.x, kt = blockIdx.y, ib = blockIdx.z;
    bool any = false, nonzero = false;
    for (int i = threadIdx.x; i < FP_BM * (FP_BN / 8); i += blockDim.x) {            // FP_BM rows x 8 chunks of 8 halves
        const int64_t row = (int64_t) qt * FP_BM + i / (FP_BN / 8), col = (int64_t) kt * FP_BN + (i % (FP_BN / 8)) * 8;
        if (row >= n_q || col >= n_kv) continue;
        const char * p = mask + row * m_nb1 + ib * m_nb3 + col * 2;
        if (col + 8 <= n_kv && ((uintptr_t) p % 16) == 0) {
            const uint4 w = __ldg((const uint4 *) p);                                   // -inf is 0xfc00, +0 is 0x0000
            const uint32_t v[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
            for (int c = 0; c < 4; ++c) { any |= (v[c] & 0xffffu) != 0xfc00u || (v[c] >> 16) != 0xfc00u; nonzero |= v[c] != 0u; }
        } else {
            for (int c = 0; c < 8 && col + c < n_kv; ++c) { const uint16_t h = ((const uint16_t *) p)[c]; any |= h != 0xfc00u; nonzero |= h != 0u; }
            nonzero |= col + 8 > n_kv;                                                  // a ragged last tile needs the bounds handling of the masked path
        }
    }
    const int r_any = __syncthreads_or(any), r_nz = __syncthreads_or(nonzero);
    if (threadIdx.x == 0) {
        if (r_any) atomicMax(out_max + ib * n_q_tiles + qt, kt + 1);
        if (r_nz) atomicMin(out_plain + ib * n_q_tiles + qt, kt);
    }
}
__global__ void k_fa_kvmax_init(int32_t * out_max, int32_t * out_plain, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out_max[i] = 0; out_plain[i] = 0x7fffffff; }
}

template <int D>
__global__ void __launch_bounds__(FP_THREADS) k_fa_prefill(const FaPArgs A) {
    constexpr int CH = D / 8;                                         // 16-byte chunks per K/V row
    constexpr int KS = D / 16;                                        // k-steps of the QK^T product
    extern __shared__ __align__(128) uint8_t fp_sm[];
    __half * Ks = (__half *) fp_sm;                                   // [2][64][D]
    __half * Vs = Ks + 2 * FP_BN * D;                                 // [2][64][D]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int qt = gridDim.x - 1 - blockIdx.x, head = blockIdx.y; const int64_t ib = blockIdx.z;      // causal: the longest tiles are scheduled first
    const int kvh = head / (int) (A.n_head / A.n_head_kv);
    const int64_t q0 = (int64_t) qt * FP_BM + warp * 16;              // first query row of this warp

    // ---- Q fragments: f32 -> f16 (the oracle rounds Q to f16), rows g and g+8 of the warp's 16
    uint32_t qf[KS][4];
    {
        const int64_t r0 = min(q0 + g, A.n_q - 1), r1 = min(q0 + g + 8, A.n_q - 1);
        const float * p0 = (const float *) (A.q + r0 * A.q_nb1 + (int64_t) head * A.q_nb2 + ib * A.q_nb3);
        const float * p1 = (const float *) (A.q + r1 * A.q_nb1 + (int64_t) head * A.q_nb2 + ib * A.q_nb3);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const float2 a = *(const float2 *) (p0 + 16 * ks + 2 * t4), b = *(const float2 *) (p1 + 16 * ks + 2 * t4);
            const float2 c = *(const float2 *) (p0 + 16 * ks + 2 * t4 + 8), d = *(const float2 *) (p1 + 16 * ks + 2 * t4 + 8);
            qf[ks][0] = fp_pack(a.x, a.y); qf[ks][1] = fp_pack(b.x, b.y); qf[ks][2] = fp_pack(c.x, c.y); qf[ks][3] = fp_pack(d.x, d.y);
        }
    }
    const char * kb = A.k + (int64_t) kvh * A.k_nb2 + (ib % A.k_ne3) * A.k_nb3;
    const char * vb = A.v + (int64_t) kvh * A.v_nb2 + (ib % A.k_ne3) * A.v_nb3;
    const int n_kv_tiles_all = (int) ((A.n_kv + FP_BN - 1) / FP_BN);
    const int n_tiles = A.kv_tiles ? min(A.kv_tiles[(ib % A.m_ne3) * gridDim.x + qt], n_kv_tiles_all) : n_kv_tiles_all;
    const int n_plain = A.kv_plain ? A.kv_plain[(ib % A.m_ne3) * gridDim.x + qt] : 0;          // tiles j < n_plain: mask is all zero

    // swizzled tile fill: row r, chunk c -> chunk c ^ (r & 7)
    auto load_tile = [&](int j, int buf) {
        __half * kd = Ks + buf * FP_BN * D, * vd = Vs + buf * FP_BN * D;
        for (int i = tid; i < FP_BN * CH; i += FP_THREADS) {
            const int r = i / CH, c = i % CH;
            const int64_t pos = (int64_t) j * FP_BN + r;
            const bool ok = pos < A.n_kv;
            const int64_t p = ok ? pos : 0;
            const uint32_t off = (uint32_t) (r * D + ((c ^ (r & 7)) * 8)) * 2;
            fp_cp16(fp_smem(kd) + off, kb + p * A.k_nb1 + c * 16, ok);
            fp_cp16(fp_smem(vd) + off, vb + p * A.v_nb1 + c * 16, ok);
        }
        fp_commit();
    };

    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f; }
    float m_run[2] = { -INFINITY, -INFINITY }, l_run[2] = { 0.0f, 0.0f };
    const int64_t row0 = q0 + g, row1 = q0 + g + 8;
    const __half * mrow0 = A.mask ? (const __half *) (A.mask + min(row0, A.n_q - 1) * A.m_nb1 + (ib % A.m_ne3) * A.m_nb3) : nullptr;
    const __half * mrow1 = A.mask ? (const __half *) (A.mask + min(row1, A.n_q - 1) * A.m_nb1 + (ib % A.m_ne3) * A.m_nb3) : nullptr;

    if (n_tiles > 0) load_tile(0, 0);
    for (int j = 0; j < n_tiles; ++j) {
        const int buf = j & 1;
        if (j + 1 < n_tiles) { load_tile(j + 1, buf ^ 1); fp_wait<1>(); } else fp_wait<0>();
        __syncthreads();
        const uint32_t ks_base = fp_smem(Ks + buf * FP_BN * D), vs_base = fp_smem(Vs + buf * FP_BN * D);

        // ---- S = Q . K^T : 8 n-blocks (8 kv each) x KS k-steps
        float s[FP_BN / 8][4];
#pragma unroll
        for (int nb = 0; nb < FP_BN / 8; ++nb) {
            s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.0f;
            const int r = nb * 8 + (lane & 7);
#pragma unroll
            for (int kp = 0; kp < KS / 2; ++kp) {                      // two k-steps per ldmatrix.x4: chunks 4kp .. 4kp+3
                uint32_t b[4];
                const int c = 4 * kp + (lane >> 3);
                fp_ldsm4(ks_base + (uint32_t) (r * D + ((c ^ (r & 7)) * 8)) * 2, b);
                fp_mma(s[nb], qf[2 * kp], b[0], b[1]);
                fp_mma(s[nb], qf[2 * kp + 1], b[2], b[3]);
            }
        }
        // ---- scale + mask, online softmax (rows g and g+8; 4 lanes share a row)
        float mx[2] = { -INFINITY, -INFINITY };
        if (j < n_plain) {
#pragma unroll
            for (int nb = 0; nb < FP_BN / 8; ++nb) {
                s[nb][0] *= A.scale; s[nb][1] *= A.scale; s[nb][2] *= A.scale; s[nb][3] *= A.scale;
                mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1])); mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
            }
        } else
#pragma unroll
        for (int nb = 0; nb < FP_BN / 8; ++nb) {
            const int64_t col = (int64_t) j * FP_BN + nb * 8 + 2 * t4;
            float m00 = 0.0f, m01 = 0.0f, m10 = 0.0f, m11 = 0.0f;
            if (mrow0) {
                if (col + 1 < A.n_kv) { const float2 a = __half22float2(*(const __half2 *) (mrow0 + col)), b = __half22float2(*(const __half2 *) (mrow1 + col)); m00 = a.x; m01 = a.y; m10 = b.x; m11 = b.y; }
                else if (col < A.n_kv) { m00 = __half2float(mrow0[col]); m10 = __half2float(mrow1[col]); }
            }
            if (col >= A.n_kv) { m00 = m10 = -INFINITY; }
            if (col + 1 >= A.n_kv) { m01 = m11 = -INFINITY; }
            s[nb][0] = m00 == -INFINITY ? -INFINITY : s[nb][0] * A.scale + m00; s[nb][1] = m01 == -INFINITY ? -INFINITY : s[nb][1] * A.scale + m01;
            s[nb][2] = m10 == -INFINITY ? -INFINITY : s[nb][2] * A.scale + m10; s[nb][3] = m11 == -INFINITY ? -INFINITY : s[nb][3] * A.scale + m11;
            mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1])); mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
        }
        float corr[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1)); mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
            const float m_new = fmaxf(m_run[h], mx[h]);
            corr[h] = m_new == -INFINITY ? 1.0f : __expf(m_run[h] - m_new);
            m_run[h] = m_new;
        }
        float sum[2] = { 0.0f, 0.0f };
        const float mb0 = m_run[0] == -INFINITY ? 0.0f : m_run[0], mb1 = m_run[1] == -INFINITY ? 0.0f : m_run[1];
#pragma unroll
        for (int nb = 0; nb < FP_BN / 8; ++nb) {
            s[nb][0] = __expf(s[nb][0] - mb0); s[nb][1] = __expf(s[nb][1] - mb0); s[nb][2] = __expf(s[nb][2] - mb1); s[nb][3] = __expf(s[nb][3] - mb1);
            sum[0] += s[nb][0] + s[nb][1]; sum[1] += s[nb][2] + s[nb][3];
        }
        l_run[0] = l_run[0] * corr[0] + sum[0]; l_run[1] = l_run[1] * corr[1] + sum[1];       // per-lane partial sums; reduced over the 4 lanes at the end
#pragma unroll
        for (int i = 0; i < D / 8; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
        // ---- O += P . V : 4 k-steps (16 kv each) x D/8 d-blocks; V fragments through ldmatrix.trans
#pragma unroll
        for (int kk = 0; kk < FP_BN / 16; ++kk) {
            const uint32_t pa[4] = { fp_pack(s[2 * kk][0], s[2 * kk][1]), fp_pack(s[2 * kk][2], s[2 * kk][3]),
                                     fp_pack(s[2 * kk + 1][0], s[2 * kk + 1][1]), fp_pack(s[2 * kk + 1][2], s[2 * kk + 1][3]) };
            const int r = 16 * kk + 8 * ((lane >> 3) & 1) + (lane & 7);
#pragma unroll
            for (int dp = 0; dp < D / 16; ++dp) {                       // two d-blocks per ldmatrix.x4.trans
                uint32_t b[4];
                const int c = 2 * dp + (lane >> 4);
                fp_ldsm4t(vs_base + (uint32_t) (r * D + ((c ^ (r & 7)) * 8)) * 2, b);
                fp_mma(o[2 * dp], pa, b[0], b[1]);
                fp_mma(o[2 * dp + 1], pa, b[2], b[3]);
            }
        }
        __syncthreads();                                                // the buffer is refilled two iterations later
    }
    // ---- normalise and store: dst[d, head, q]  (row g: o[.][0..1], row g+8: o[.][2..3])
#pragma unroll
    for (int h = 0; h < 2; ++h) { l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1); l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2); }
    const float inv0 = l_run[0] == 0.0f ? 0.0f : 1.0f / l_run[0], inv1 = l_run[1] == 0.0f ? 0.0f : 1.0f / l_run[1];
    float * d0 = (float *) (A.dst + (int64_t) head * A.d_nb1 + row0 * A.d_nb2 + ib * A.d_nb3);
    float * d1 = (float *) (A.dst + (int64_t) head * A.d_nb1 + row1 * A.d_nb2 + ib * A.d_nb3);
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        if (row0 < A.n_q) *(float2 *) (d0 + 8 * i + 2 * t4) = make_float2(o[i][0] * inv0, o[i][1] * inv0);
        if (row1 < A.n_q) *(float2 *) (d1 + 8 * i + 2 * t4) = make_float2(o[i][2] * inv1, o[i][3] * inv1);
    }
}

// ---------------------------------------------------------------------------------------------------------------- host side
// the tcgen05 / TMEM / TMA kernel (fa_tc.cu) takes head size 128 with an F16 K/V cache; this file's mma.sync kernel keeps head size 64 and odd layouts
bool fa_tc_supported(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask, const b200_tensor * dst);
int  fa_tc(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask, const b200_tensor * dst, float scale,
           const int32_t * kv_tiles, const int32_t * kv_plain, void * tiles, cudaStream_t st);

bool fa_prefill_supported(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask, const b200_tensor * dst) {
    if (q->ne[1] < 16) return false;                                   // few query tokens: the split-KV decode kernel (flash_attn.cu)
    if (q->ne[0] != 128 && q->ne[0] != 64) return false;
    if ((uintptr_t) q->data % 8 || q->nb[1] % 8 || q->nb[2] % 8 || q->nb[3] % 8) return false;                 // float2 loads of Q
    if ((uintptr_t) dst->data % 8 || dst->nb[1] % 8 || dst->nb[2] % 8 || dst->nb[3] % 8) return false;
    if (mask && ((uintptr_t) mask->data % 4 || mask->nb[1] % 4 || mask->nb[3] % 4)) return false;
    (void) k; (void) v;
    return true;
}
size_t fa_prefill_scratch_bytes(const b200_tensor * q, const b200_tensor * mask_or_null, int64_t m_ne3) {
    (void) mask_or_null;
    return (size_t) ((q->ne[1] + FP_BM - 1) / FP_BM) * (size_t) (m_ne3 > 0 ? m_ne3 : 1) * 8 + 16;       // kv_tiles + kv_plain
}

int fa_prefill(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask, const b200_tensor * dst, float scale,
               void * scratch, cudaStream_t st, void * tiles) {
    static smem_mask_t done128{0}, done64{0};
    B200_CUDA_TRY(ensure_dyn_smem(k_fa_prefill<128>, 4 * FP_BN * 128 * 2, done128));
    B200_CUDA_TRY(ensure_dyn_smem(k_fa_prefill<64>, 4 * FP_BN * 64 * 2, done64));
    FaPArgs A = {};
    A.q = (const char *) q->data; A.k = (const char *) k->data; A.v = (const char *) v->data; A.mask = mask ? (const char *) mask->data : nullptr; A.dst = (char *) dst->data;
    A.q_nb1 = q->nb[1]; A.q_nb2 = q->nb[2]; A.q_nb3 = q->nb[3]; A.k_nb1 = k->nb[1]; A.k_nb2 = k->nb[2]; A.k_nb3 = k->nb[3];
    A.v_nb1 = v->nb[1]; A.v_nb2 = v->nb[2]; A.v_nb3 = v->nb[3]; A.d_nb1 = dst->nb[1]; A.d_nb2 = dst->nb[2]; A.d_nb3 = dst->nb[3];
    A.n_q = q->ne[1]; A.n_kv = k->ne[1]; A.n_head = q->ne[2]; A.n_head_kv = k->ne[2]; A.k_ne3 = k->ne[3]; A.scale = scale; A.m_ne3 = 1;
    const int n_q_tiles = (int) ((A.n_q + FP_BM - 1) / FP_BM);
    if (mask) {
        A.m_nb1 = mask->nb[1]; A.m_nb3 = mask->nb[3]; A.m_ne3 = mask->ne[3];
        if (scratch) {
            const int n_kv_tiles = (int) ((A.n_kv + FP_BN - 1) / FP_BN);
            const int n_ent = n_q_tiles * (int) A.m_ne3;
            int32_t * kmax = (int32_t *) scratch, * kplain = kmax + n_ent;
            k_fa_kvmax_init<<<(n_ent + 127) / 128, 128, 0, st>>>(kmax, kplain, n_ent);
            k_fa_kvmax<<<dim3((unsigned) n_q_tiles, (unsigned) n_kv_tiles, (unsigned) A.m_ne3), 128, 0, st>>>(A.mask, A.m_nb1, A.m_nb3, A.n_q, A.n_kv, kmax, kplain, n_q_tiles);
            B200_LAUNCH_CHECK();
            A.kv_tiles = kmax; A.kv_plain = kplain;
        }
    }
    if (fa_tc_supported(q, k, v, mask, dst)) return fa_tc(q, k, v, mask, dst, scale, A.kv_tiles, A.kv_plain, tiles, st);
    if (tiles) return B200_ERR_UNSUPPORTED;                            // only the tcgen05 kernel writes activation tiles
    const dim3 grid((unsigned) n_q_tiles, (unsigned) A.n_head, (unsigned) q->ne[3]);
    if (q->ne[0] == 128) k_fa_prefill<128><<<grid, FP_THREADS, 4 * FP_BN * 128 * 2, st>>>(A);
    else                 k_fa_prefill<64><<<grid, FP_THREADS, 4 * FP_BN * 64 * 2, st>>>(A);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

} // namespace b200
