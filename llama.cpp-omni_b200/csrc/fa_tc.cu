// fa_tc.cu — GGML_OP_FLASH_ATTN_EXT for prefill on the 5th-generation tensor cores: S = Q.K^T and O_j = P.V_j are tcgen05.mma (kind::f16, M = 128,
// N = 128, K = 16) with the accumulators in TMEM, K / V tiles arrive by 4-D tiled TMA with the 128-byte swizzle, the online softmax runs on
// 256 threads, two per query row (a row = one TMEM lane).
//
// Replaces flash_attn_ext_f16 (ggml-cuda/fattn-mma-f16.cuh:1246, dispatch fattn.cu:295) for head size 128 with an F16 K/V cache — the reference's
// kernel (and round 1's k_fa_prefill) is mma.sync + cp.async, the previous generation's data path.  Oracle: ggml-cpu/ops.cpp:7912-8148 (Q rounded to
// F16, F32 scores, F32 softmax; the CPU accumulates V in F16, we in F32 — the op's own bar is NMSE <= 5e-4, tests/test-backend-ops.cpp:5085).
//
// One CTA = 128 query rows of one head.  12 warps:
//   warp 0      TMA producer: per KV tile (128 positions) two boxes {64 dims, 128 positions} of K and two of V -> 2-stage ring (4 x 16 KB per stage).
//               The SAME box shape serves both operands: K is the B operand of S = Q.K^T in K-major form (dims contiguous), V the B operand of
//               O = P.V in MN-major form (dims contiguous, positions = the reduction index) — only the shared-memory descriptors differ.
//   warp 1      MMA issuer (one lane): S_j = Q.K_j^T into one of two 128-column TMEM buffers (so QK of tile j+1 overlaps the softmax of tile j),
//               then O_j = P_j.V_j into a third 128-column buffer; tcgen05.commit publishes each to the softmax warps / frees the stage.
//   warps 4-11  softmax: two threads per query row (TMEM lane = row; warps w and w + 4 share a lane quarter and split the row's 128 score columns and
//               its 128 output dims in halves).  Two passes over the scores in TMEM (tcgen05.ld 32x32b.x32): running max (exchanged between the two
//               halves through shared memory), then p = exp2(s * scale * log2e - m) -> F16 P tile in shared memory (canonical K-major UMMA layout,
//               16-byte stores) for the P.V product.  acc = acc * corr + O_j is applied one tile late (while the tensor core already works on the
//               next S), with the 64 output accumulators of the thread's half row in registers.
//   warps 2-3   help convert the Q tile (F32 in global memory) to the F16 UMMA tile, then idle.
// Mask handling follows k_fa_prefill: k_fa_kvmax (fa_prefill.cu) tells, per query tile, how many KV tiles hold an unmasked position (a causal mask
// skips everything right of the diagonal) and how many leading tiles are mask-free (no mask loads at all).
// FLOPs per CTA and KV tile = 4 * 128^3.
#include "common.cuh"
#include <cuda.h>
#include <math.h>

namespace b200 {

constexpr int FT_BM = 128, FT_BN = 128, FT_D = 128, FT_THREADS = 384, FT_STAGES = 2;
constexpr int FT_TILE = FT_BM * FT_D * 2;                                 // 32 KB: a 128 x 128 F16 operand tile
constexpr int FT_SMEM = 2 * FT_TILE /* Q, P */ + FT_STAGES * 2 * FT_TILE /* K, V */ + 1024 /* alignment */ + 256 /* barriers */ + 3 * 2 * FT_BM * 4 /* row max (double-buffered) / row sum exchange */;
static_assert(FT_SMEM <= 227 * 1024, "k_fa_tc shared memory");
constexpr uint32_t FT_LBO = (FT_BM / 8) * 128, FT_SBO = 128;             // canonical K-major, no swizzle: 8 x 16-byte core matrices, K direction 2 KB apart

struct FaTcArgs {
    const char * q; const char * mask; char * dst;
    int64_t q_nb1, q_nb2, q_nb3, m_nb1, m_nb3, d_nb1, d_nb2, d_nb3;
    int64_t n_q, n_kv, n_head, n_head_kv, k_ne3, m_ne3;
    int kc_pos, kc_head, kc_b;                                             // which TMA coordinate (1..3) carries the position / kv head / batch index
    float scale;
    const int32_t * kv_tiles; const int32_t * kv_plain;                    // per (mask batch, query tile), in units of 64 positions (k_fa_kvmax); null = visit all / mask all
    uint8_t * tiles;                                                       // optional: the [n_q, n_head * 128] result as F16 activation tiles for the wo MUL_MAT (act_tile_off) instead of F32
};

// ---------------------------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t ft_smem(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void ft_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void ft_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void ft_mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void ft_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void ft_tma_4d(uint32_t dst, const CUtensorMap * map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets in 16-byte units, version 1, layout type
__device__ __forceinline__ uint64_t ft_desc_k_noswz(uint32_t saddr) {       // K-major, SWIZZLE_NONE: LBO = K direction, SBO = 8-row groups
    return (uint64_t) ((saddr & 0x3ffff) >> 4) | ((uint64_t) (FT_LBO >> 4) << 16) | ((uint64_t) (FT_SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t ft_desc_k_sw128(uint32_t saddr) {       // K-major, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart (LBO unused)
    return (uint64_t) ((saddr & 0x3ffff) >> 4) | ((uint64_t) 1 << 16) | ((uint64_t) (1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t ft_desc_mn_sw128(uint32_t saddr) {      // MN-major, SWIZZLE_128B: 64-element MN atoms FT_TILE/2 apart (LBO), 8 K-rows 1024 B apart (SBO)
    return (uint64_t) ((saddr & 0x3ffff) >> 4) | ((uint64_t) ((FT_TILE / 2) >> 4) << 16) | ((uint64_t) (1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = F32 (bit 4), A = B = F16, A K-major, B K-major or MN-major (bit 16), N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t ft_idesc(bool b_mn_major) { return (1u << 4) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t) (FT_BN >> 3) << 17) | ((uint32_t) (FT_BM >> 4) << 24); }
__device__ __forceinline__ void ft_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ft_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void ft_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ft_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ft_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                 "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                   "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ft_ld32_issue(uint32_t taddr, float * v) {      // 32 consecutive columns of this thread's TMEM lane; complete after ft_ld_wait()
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                 "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),
                   "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]),
                   "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void ft_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ft_sts16(uint32_t a, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ float ft_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }      // 2^x, one MUFU op; 2^-inf = +0
__device__ __forceinline__ uint32_t ft_pack(float lo, float hi) { const __half2 h = __floats2half2_rn(lo, hi); return *(const uint32_t *) &h; }

// ---------------------------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(FT_THREADS, 1) k_fa_tc(const FaTcArgs A, const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap) {
    extern __shared__ __align__(1024) uint8_t smem_ft[];
    const uint32_t sbase = (ft_smem(smem_ft) + 1023u) & ~1023u;                // SWIZZLE_128B tiles sit on 1024-byte boundaries
    const uint32_t q_s = sbase, p_s = sbase + FT_TILE, kv_s = sbase + 2 * FT_TILE;
    const uint32_t bars = kv_s + FT_STAGES * 2 * FT_TILE;
    const uint32_t kv_full = bars, kv_empty = bars + 16, s_full = bars + 32, s_empty = bars + 48, p_full = bars + 64, p_empty = bars + 72, o_full = bars + 80, o_empty = bars + 88,
                   tmem_slot = bars + 96, xch = bars + 256;                                      // xch: float [2 buffers][2 halves][128 rows]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // heaviest query tiles first (a causal mask gives tile i about i + 1 KV tiles)
    // longest-processing-time-first over the WHOLE grid: blocks are dispatched in blockIdx.x-fastest order, so the head index is x and the query tile y, last tile
    // (the causal row block with the most KV tiles) first — every head's heaviest tile starts before any light one, and the last wave holds only 1-tile blocks
    const int n_qt = (int) gridDim.y, qt = n_qt - 1 - (int) blockIdx.y, head = blockIdx.x, ib = blockIdx.z;
    const int kvh = head / (int) (A.n_head / A.n_head_kv), ibk = (int) (ib % A.k_ne3), ibm = (int) (ib % A.m_ne3);
    const int64_t q0 = (int64_t) qt * FT_BM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < FT_STAGES; ++s) { ft_mbar_init(kv_full + 8 * s, 1); ft_mbar_init(kv_empty + 8 * s, 1); ft_mbar_init(s_full + 8 * s, 1); ft_mbar_init(s_empty + 8 * s, 256); }
        ft_mbar_init(p_full, 256); ft_mbar_init(p_empty, 1); ft_mbar_init(o_full, 1); ft_mbar_init(o_empty, 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- Q tile: F32 rows in global memory -> F16, canonical K-major UMMA layout (row r, dims 8c..8c+7 at c * LBO + (r / 8) * SBO + (r % 8) * 16) ----------
    if (threadIdx.x < 256) {
        const int r = threadIdx.x & 127, hf = threadIdx.x >> 7;                 // thread = (row, half of the 128 dims)
        const bool live = q0 + r < A.n_q;
        const float * src = (const float *) (A.q + (q0 + r) * A.q_nb1 + (int64_t) head * A.q_nb2 + (int64_t) ib * A.q_nb3) + hf * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float4 a = make_float4(0, 0, 0, 0), b = a;
            if (live) { a = *(const float4 *) (src + c * 8); b = *(const float4 *) (src + c * 8 + 4); }
            ft_sts16(q_s + (uint32_t) (hf * 8 + c) * FT_LBO + (uint32_t) (r >> 3) * FT_SBO + (uint32_t) (r & 7) * 16,
                     make_uint4(ft_pack(a.x, a.y), ft_pack(a.z, a.w), ft_pack(b.x, b.y), ft_pack(b.z, b.w)));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    ft_fence_before();
    __syncthreads();
    ft_fence_after();
    uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

    // KV tiles to visit / leading mask-free tiles, in units of FT_BN positions (k_fa_kvmax counts 64-position tiles of a 128-row query tile)
    const int n_kv_tiles = (int) ((A.n_kv + FT_BN - 1) / FT_BN);
    int n_visit = n_kv_tiles, n_plain = A.mask ? 0 : n_kv_tiles;
    if (A.kv_tiles) {
        const int ent = ibm * n_qt + qt;
        n_visit = min(n_kv_tiles, (A.kv_tiles[ent] + 1) / 2);
        n_plain = min(n_visit, A.kv_plain[ent] == 0x7fffffff ? n_visit : A.kv_plain[ent] / 2);
    }
    if ((int64_t) n_plain * FT_BN > A.n_kv) n_plain = (int) (A.n_kv / FT_BN);                   // a ragged last tile takes the bounds-checked path

    if (warp == 0) {
        if (lane == 0) {
            for (int j = 0; j < n_visit; ++j) {
                const uint32_t s = j % FT_STAGES;
                ft_mbar_wait(kv_empty + 8 * s, ((j / FT_STAGES) & 1) ^ 1);
                ft_mbar_expect_tx(kv_full + 8 * s, 2 * FT_TILE);
                const uint32_t k_s = kv_s + s * 2 * FT_TILE, v_s = k_s + FT_TILE;
                int c[4] = { 0, 0, 0, 0 };
                c[A.kc_pos] = j * FT_BN; c[A.kc_head] = kvh; c[A.kc_b] = ibk;
                ft_tma_4d(k_s,               &kmap, 0,  c[1], c[2], c[3], kv_full + 8 * s);
                ft_tma_4d(k_s + FT_TILE / 2, &kmap, 64, c[1], c[2], c[3], kv_full + 8 * s);
                ft_tma_4d(v_s,               &vmap, 0,  c[1], c[2], c[3], kv_full + 8 * s);
                ft_tma_4d(v_s + FT_TILE / 2, &vmap, 64, c[1], c[2], c[3], kv_full + 8 * s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_qk = ft_idesc(false), idesc_pv = ft_idesc(true);
            auto issue_qk = [&](int j) {
                const uint32_t s = j % FT_STAGES, b = j & 1;
                ft_mbar_wait(kv_full + 8 * s, (j / FT_STAGES) & 1);
                ft_mbar_wait(s_empty + 8 * b, ((j >> 1) & 1) ^ 1);
                ft_fence_after();
                const uint32_t k_s = kv_s + s * 2 * FT_TILE;
#pragma unroll
                for (int kk = 0; kk < FT_D / 16; ++kk)                          // k-step = 16 dims: 2 core-matrix columns of Q; 32 bytes inside a 64-dim K box
                    ft_mma(tmem + b * FT_BN, ft_desc_k_noswz(q_s + kk * 2 * FT_LBO), ft_desc_k_sw128(k_s + (kk >> 2) * (FT_TILE / 2) + (kk & 3) * 32), idesc_qk, kk != 0);
                ft_commit(s_full + 8 * b);
            };
            if (n_visit > 0) issue_qk(0);
            for (int j = 0; j < n_visit; ++j) {
                if (j + 1 < n_visit) issue_qk(j + 1);                           // overlaps the softmax of tile j
                const uint32_t s = j % FT_STAGES;
                ft_mbar_wait(p_full, j & 1);
                ft_mbar_wait(o_empty, (j & 1) ^ 1);
                ft_fence_after();
                const uint32_t v_s = kv_s + s * 2 * FT_TILE + FT_TILE;
#pragma unroll
                for (int kk = 0; kk < FT_BN / 16; ++kk)                         // k-step = 16 positions: 2 core-matrix columns of P; 16 rows (2 KB) of the V boxes
                    ft_mma(tmem + 2 * FT_BN, ft_desc_k_noswz(p_s + kk * 2 * FT_LBO), ft_desc_mn_sw128(v_s + kk * 16 * 128), idesc_pv, kk != 0);
                ft_commit(o_full);
                ft_commit(p_empty);
                ft_commit(kv_empty + 8 * s);
            }
        }
    } else if (warp >= 4) {
        const int quarter = warp & 3, hf = (warp - 4) >> 2, r = quarter * 32 + lane;      // query row = TMEM lane; hf = which 64 score columns / output dims
        const uint32_t lane_addr = tmem + ((uint32_t) (quarter * 32) << 16);
        const bool live = q0 + r < A.n_q;
        const char * mrow = A.mask ? A.mask + (q0 + (live ? r : 0)) * A.m_nb1 + (int64_t) ibm * A.m_nb3 : nullptr;
        const float sl2 = A.scale * 1.44269504088896f;                          // scores live in the log2 domain: p = exp2(s * scale * log2e + mask * log2e - m)
        float acc[FT_D / 2];
#pragma unroll
        for (int i = 0; i < FT_D / 2; ++i) acc[i] = 0.0f;
        float m_run = -INFINITY, l_run = 0.0f, corr_prev = 1.0f;
        // mask (log2 domain) applied to this thread's 64 scores of tile j; afterwards t holds log2-domain scores, -inf where masked / out of range
        auto apply_mask = [&](int j, float (&t)[64]) {
            const int64_t col0 = (int64_t) j * FT_BN + hf * 64;
            const bool vec = mrow && ((A.m_nb1 | A.m_nb3 | (int64_t) (uintptr_t) A.mask) % 16) == 0;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                uint4 mk = make_uint4(0, 0, 0, 0);
                if (vec && col0 + g * 8 + 8 <= A.n_kv) mk = __ldg((const uint4 *) (mrow + (col0 + g * 8) * 2));
                else if (mrow) { __half hh[8]; for (int i = 0; i < 8; ++i) hh[i] = col0 + g * 8 + i < A.n_kv ? ((const __half *) mrow)[col0 + g * 8 + i] : __float2half(0.0f); mk = *(const uint4 *) hh; }
                const __half2 * h2 = (const __half2 *) &mk;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(h2[i]);
                    const int e = g * 8 + 2 * i;
                    t[e]     = (col0 + e     >= A.n_kv || f.x == -INFINITY) ? -INFINITY : fmaf(t[e],     sl2, f.x * 1.44269504088896f);
                    t[e + 1] = (col0 + e + 1 >= A.n_kv || f.y == -INFINITY) ? -INFINITY : fmaf(t[e + 1], sl2, f.y * 1.44269504088896f);
                }
            }
        };
        auto accumulate = [&](int j, float corr) {                                  // acc = acc * corr + O_j (this thread's 64 dims)
            ft_mbar_wait(o_full, j & 1);
            ft_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float ov[32];
                ft_ld32_issue(lane_addr + 2 * FT_BN + hf * 64 + c * 32, ov);
                ft_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[c * 32 + i] = fmaf(acc[c * 32 + i], corr, ov[i]);
            }
            ft_fence_before();
            ft_mbar_arrive(o_empty);
        };
        for (int j = 0; j < n_visit; ++j) {
            const uint32_t b = j & 1;
            const bool plain = j < n_plain && A.scale > 0.0f;
            ft_mbar_wait(s_full + 8 * b, (j >> 1) & 1);
            ft_fence_after();
            // ---- this thread's 64 scores: ONE round trip to TMEM, kept in registers for both passes; the S buffer is free again right away ----------
            float t[64];
            ft_ld32_issue(lane_addr + b * FT_BN + hf * 64, t);
            ft_ld32_issue(lane_addr + b * FT_BN + hf * 64 + 32, t + 32);
            ft_ld_wait();
            ft_fence_before();
            ft_mbar_arrive(s_empty + 8 * b);                                    // Q.K of tile j + 2 may overwrite it
            if (!plain) apply_mask(j, t);
            // ---- pass 1: maximum (4 independent chains), exchanged with the thread that owns the other half of the row (64-thread named barrier) -----
            float mx4[4] = { -INFINITY, -INFINITY, -INFINITY, -INFINITY };
#pragma unroll
            for (int i = 0; i < 64; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], t[i]);
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * (plain ? sl2 : 1.0f);     // mask-free tiles keep RAW scores: the positive scale is folded in here and below
            asm volatile("st.shared.f32 [%0], %1;" :: "r"(xch + ((b * 2 + hf) * FT_BM + r) * 4), "f"(mx) : "memory");
            asm volatile("bar.sync %0, 64;" :: "r"(1 + quarter) : "memory");    // warps w and w + 4 only
            float mo; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mo) : "r"(xch + ((b * 2 + (hf ^ 1)) * FT_BM + r) * 4));
            const float m_new = fmaxf(m_run, fmaxf(mx, mo));
            const float m_safe = m_new == -INFINITY ? 0.0f : m_new;              // a row with nothing unmasked yet: every p below is 2^-inf = 0
            const float corr = ft_ex2(m_run - m_safe);
            // ---- the previous tile's P.V result is folded in while the tensor core is busy with this tile's neighbours --------------------------
            if (j > 0) accumulate(j - 1, corr_prev);                             // (also implies that P.V of tile j - 1 has consumed the P buffer)
            // ---- pass 2: p = exp2(t - m) -> F16 P tile (A operand of P.V) ------------------------------------------------------------------------
            float ls4[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const float p = plain ? ft_ex2(fmaf(t[i], sl2, -m_safe)) : ft_ex2(t[i] - m_safe);
                t[i] = p; ls4[i & 3] += p;
            }
#pragma unroll
            for (int g = 0; g < 8; ++g)                                         // 8 positions = one 16-byte core-matrix row: K chunk 8 hf + g of row r
                ft_sts16(p_s + (uint32_t) (hf * 8 + g) * FT_LBO + (uint32_t) (r >> 3) * FT_SBO + (uint32_t) (r & 7) * 16,
                         make_uint4(ft_pack(t[8 * g], t[8 * g + 1]), ft_pack(t[8 * g + 2], t[8 * g + 3]), ft_pack(t[8 * g + 4], t[8 * g + 5]), ft_pack(t[8 * g + 6], t[8 * g + 7])));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores of P before the tensor core (async proxy) reads them
            ft_mbar_arrive(p_full);
            l_run = l_run * corr + (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]); m_run = m_new; corr_prev = corr;
        }
        if (n_visit > 0) accumulate(n_visit - 1, corr_prev);
        // combine the two halves' partial row sums (same running maximum), normalise, store this thread's 64 dims
        asm volatile("st.shared.f32 [%0], %1;" :: "r"(xch + ((4 + hf) * FT_BM + r) * 4), "f"(l_run) : "memory");
        asm volatile("bar.sync %0, 64;" :: "r"(1 + quarter) : "memory");
        float lo; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lo) : "r"(xch + ((4 + (hf ^ 1)) * FT_BM + r) * 4));
        const float l_tot = l_run + lo;
        if (A.tiles) {                                                          // straight into the next MUL_MAT's operand layout (rows past n_q: zeros)
            const float inv = (l_tot == 0.0f || !live) ? 0.0f : 1.0f / l_tot;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8)
                *(uint4 *) (A.tiles + act_tile_off(q0 + r, (int64_t) head * FT_D + hf * 64 + c8 * 8, A.n_head * FT_D)) =
                    make_uint4(ft_pack(acc[8 * c8] * inv, acc[8 * c8 + 1] * inv), ft_pack(acc[8 * c8 + 2] * inv, acc[8 * c8 + 3] * inv),
                               ft_pack(acc[8 * c8 + 4] * inv, acc[8 * c8 + 5] * inv), ft_pack(acc[8 * c8 + 6] * inv, acc[8 * c8 + 7] * inv));
        } else if (live) {
            const float inv = l_tot == 0.0f ? 0.0f : 1.0f / l_tot;
            float * out = (float *) (A.dst + (int64_t) head * A.d_nb1 + (q0 + r) * A.d_nb2 + (int64_t) ib * A.d_nb3) + hf * 64;
#pragma unroll
            for (int i = 0; i < FT_D / 2; i += 4) *(float4 *) (out + i) = make_float4(acc[i] * inv, acc[i + 1] * inv, acc[i + 2] * inv, acc[i + 3] * inv);
        }
    }
    ft_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------- host side
typedef CUresult (*ft_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static ft_encode_fn ft_encoder() {
    static ft_encode_fn fn = [] {
        void * p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (ft_encode_fn) p;
    }();
    return fn;
}

// K / V as a 4-D F16 tensor map: dim 0 = the 128 head dims; dims 1..3 = (position, kv head, batch) ordered by ascending byte stride; box {64, 128 positions, 1, 1}
static bool ft_make_map(CUtensorMap * map, const b200_tensor * t, int & c_pos, int & c_head, int & c_b) {
    int order[3] = { 1, 2, 3 };                                            // ggml dims of K / V: ne[1] positions, ne[2] kv heads, ne[3] batch
    for (int a = 0; a < 3; ++a) for (int b = a + 1; b < 3; ++b) if (t->nb[order[b]] < t->nb[order[a]]) { const int x = order[a]; order[a] = order[b]; order[b] = x; }
    cuuint64_t gdim[4] = { (cuuint64_t) t->ne[0], 0, 0, 0 }, gstr[3];
    cuuint32_t box[4] = { 64, 1, 1, 1 }, estr[4] = { 1, 1, 1, 1 };
    for (int i = 0; i < 3; ++i) {
        gdim[i + 1] = (cuuint64_t) t->ne[order[i]]; gstr[i] = (cuuint64_t) t->nb[order[i]];
        if (order[i] == 1) { c_pos = i + 1; box[i + 1] = FT_BN; } else if (order[i] == 2) c_head = i + 1; else c_b = i + 1;
        if (gstr[i] % 16 || gstr[i] == 0) return false;
    }
    return ft_encoder()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, t->data, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool fa_tc_supported(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask, const b200_tensor * dst) {
    static const bool off = getenv("B200_DISABLE_FA_TC") && atoi(getenv("B200_DISABLE_FA_TC")) != 0;
    if (off || !ft_encoder()) return false;
    if (q->ne[0] != FT_D || q->ne[1] < 64) return false;                              // head size 128; enough query rows to fill the 128-row tiles
    if (k->type != B200_F16 || v->type != B200_F16 || k->nb[0] != 2 || v->nb[0] != 2) return false;
    if ((uintptr_t) k->data % 16 || (uintptr_t) v->data % 16) return false;
    for (int i = 1; i < 4; ++i) if (k->nb[i] % 16 || v->nb[i] % 16 || k->nb[i] == 0 || v->nb[i] == 0) return false;
    if ((uintptr_t) q->data % 16 || q->nb[1] % 16 || q->nb[2] % 16 || q->nb[3] % 16) return false;          // float4 loads of Q rows
    if ((uintptr_t) dst->data % 16 || dst->nb[1] % 16 || dst->nb[2] % 16 || dst->nb[3] % 16) return false;
    if (mask && mask->type != B200_F16) return false;
    if (q->ne[2] > 65535 || q->ne[3] > 65535) return false;
    return true;
}

int fa_tc(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask, const b200_tensor * dst, float scale,
          const int32_t * kv_tiles, const int32_t * kv_plain, void * tiles, cudaStream_t st) {
    static smem_mask_t done{0};
    B200_CUDA_TRY(ensure_dyn_smem(k_fa_tc, FT_SMEM, done));
    FaTcArgs A = {};
    CUtensorMap kmap, vmap;
    int cp = 1, ch = 2, cb = 3, vp = 1, vh = 2, vb = 3;
    if (!ft_make_map(&kmap, k, cp, ch, cb) || !ft_make_map(&vmap, v, vp, vh, vb) || cp != vp || ch != vh || cb != vb) return B200_ERR_UNSUPPORTED;
    A.kc_pos = cp; A.kc_head = ch; A.kc_b = cb;
    A.q = (const char *) q->data; A.mask = mask ? (const char *) mask->data : nullptr; A.dst = (char *) dst->data;
    A.q_nb1 = q->nb[1]; A.q_nb2 = q->nb[2]; A.q_nb3 = q->nb[3]; A.d_nb1 = dst->nb[1]; A.d_nb2 = dst->nb[2]; A.d_nb3 = dst->nb[3];
    A.n_q = q->ne[1]; A.n_kv = k->ne[1]; A.n_head = q->ne[2]; A.n_head_kv = k->ne[2]; A.k_ne3 = k->ne[3]; A.scale = scale; A.m_ne3 = 1;
    if (mask) { A.m_nb1 = mask->nb[1]; A.m_nb3 = mask->nb[3]; A.m_ne3 = mask->ne[3]; }
    A.kv_tiles = kv_tiles; A.kv_plain = kv_plain; A.tiles = (uint8_t *) tiles;
    const dim3 grid((unsigned) A.n_head, (unsigned) ((A.n_q + FT_BM - 1) / FT_BM), (unsigned) q->ne[3]);
    k_fa_tc<<<grid, FT_THREADS, FT_SMEM, st>>>(A, kmap, vmap);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

} // namespace b200
