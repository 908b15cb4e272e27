// stream_decode.cu — the B200 decode engine: ONE persistent kernel (one CTA per SM) that streams quantised weight rows through
// TMA (cp.async.bulk) shared-memory rings and executes a whole sequence of decode "phases" with grid barriers between them.
//
// Why (SURVEY.md §7 hard parts): a Qwen3-8B Q4_K_M token reads 4.67 GB of weights -> 0.71 ms at the measured 6.54 TB/s, but the
// reference needs 868 node launches + 253 quantize launches per token; kernel boundaries, not bytes, dominate.  Here
//   * each of the 12 warps of a CTA owns a private 3-slot ring and is its own producer: while it computes the unit in one slot its
//     elected lane has two more 4.6 KB bulk copies in flight (12 x 2 x 4.6 KB = 110 KB per SM, independent of occupancy), and when a
//     phase runs out of rows it prefetches the first units of the NEXT matvec phase before entering the grid barrier — weights never
//     depend on activations, so HBM keeps streaming through barriers, prologues and the attention phase;
//   * the arithmetic is the reference CPU backend's (q8_K / q8_0 activations, dp4a sub-block dots, one f32 multiply per block:
//     ggml-cpu/quants.c:115-149, 305-333, 550-758), computed out of shared memory, activation fragments held in registers;
//   * phase prologues (sum of partials -> RMS_NORM * weight -> q8 record, or plain quantisation) are recomputed by every CTA into its own
//     shared memory instead of being separate launches; epilogues fuse the residual ADD and SWIGLU; reductions longer than 4096
//     (ffn_down) are K-split across CTA groups and summed, in a fixed order, by the next prologue;
//   * a split-KV attention phase (same arithmetic as flash_attn.cu) runs on the same CTAs while wo's first rows are already resident.
// Replaces, for batch-1 decode, mul_mat_vec_q + quantize_q8_1 + rms_norm_f32 + rope_neox + k_set_rows + flash_attn_ext_vec +
// unary_gated_op_kernel + k_bin_bcast of the reference (ggml-cuda/mmvq.cu:139-227, quantize.cu:4-48, norm.cu:107-185,
// rope.cu:83-123, set-rows.cu:264, fattn-vec.cuh:19, unary.cu:208-228, binbcast.cu:395-443).
#include "quant_dev.cuh"
#include "stream_decode.cuh"
#include <math.h>
#include <mutex>
#include <stdlib.h>

namespace b200 {

// ---------------------------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void * src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 lds16(uint32_t a) { uint4 r; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a)); return r; }
__device__ __forceinline__ uint32_t lds4(uint32_t a) { uint32_t r; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a)); return r; }
__device__ __forceinline__ int lds_s16(uint32_t a) { int r; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(r) : "r"(a)); return r; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t r; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(a)); return r; }
__device__ __forceinline__ float lds_f32(uint32_t a) { float r; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a)); return r; }
__device__ __forceinline__ unsigned long long globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// 32 bytes in one request (LDG.E.256, sm_100): p must be 32-byte aligned; through L2 (.cg) — the data was just written by other SMs
__device__ __forceinline__ void ld256_cg(const float * p, float4 & lo, float4 & hi) {
    asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w) : "l"(p));
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned * p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// ---------------------------------------------------------------------------------------------------------------- geometry
// A matvec phase hands CTA c a K-slice (c % ksplit) and the rows [T*g/G, T*(g+1)/G) (g = c / ksplit, G = gridDim / ksplit) of the
// concatenated row space of its matrices (T rows in total; for SWIGLU the gate/up row PAIRS).  Those rows are cut into UNITS of rpu (1 or
// 2) consecutive rows of one matrix; unit u of the phase belongs to warp u % SD_WARPS.  Everything is derived arithmetically from u.
__device__ __forceinline__ void seg_build(SegTab & S, const SdPhase & P, int cta, int ncta) {
    const bool sw = P.epilogue == SD_EPI_SWIGLU;
    const int nmat = sw ? 1 : P.n_mat, ks = P.ksplit;
    const int grp = cta / ks, ngrp = ncta / ks;
    S.kpart = cta % ks; S.nm = sw ? 2 : 1; S.ksplit = ks;
    uint32_t T = 0;                                                       // T * ngrp < 2^32 (checked on the host): 32-bit divisions only
    for (int j = 0; j < nmat; ++j) T += (uint32_t) P.mat[j].rows;
    const int r0 = grp < ngrp ? (int) (T * (uint32_t) grp / (uint32_t) ngrp) : 0, r1 = grp < ngrp ? (int) (T * (uint32_t) (grp + 1) / (uint32_t) ngrp) : 0;   // leftover CTAs idle
    int mb = 0; S.upre[0] = 0;
    for (int j = 0; j < 3; ++j) {
        if (j < nmat) {
            const SdMat & M = P.mat[j];
            const int lo = max(r0, mb), hi = min(r1, mb + M.rows), first = lo - mb;
            S.nrows[j] = max(0, hi - lo);
            S.rbp[j] = M.row_bytes_p; S.rbd[j] = M.row_bytes_d; S.sub_p[j] = M.row_bytes_p / ks; S.sub_d[j] = M.row_bytes_d / ks; S.type[j] = M.type;
            const int unit_row = (S.sub_p[j] + S.sub_d[j]) * (sw ? 2 : 1);
            S.rpu[j] = (P.act_group == 256 && unit_row * 2 <= SD_SLOT_BYTES) ? 2 : 1;
            S.pay[j] = M.payload + (int64_t) first * M.row_bytes_p + (int64_t) S.kpart * S.sub_p[j];
            S.dpl[j] = M.dplane ? M.dplane + (int64_t) first * M.row_bytes_d + (int64_t) S.kpart * S.sub_d[j] : nullptr;
            S.y[j] = M.y + (int64_t) S.kpart * P.y_part_stride + first;
            S.resid[j] = M.residual ? M.residual + first : nullptr;
            if (sw) {
                S.pay2 = P.mat[1].payload + (int64_t) first * M.row_bytes_p;
                S.dpl2 = P.mat[1].dplane ? P.mat[1].dplane + (int64_t) first * M.row_bytes_d : nullptr;
            }
            mb += M.rows;
        } else { S.nrows[j] = 0; S.rpu[j] = 1; S.sub_p[j] = 0; S.sub_d[j] = 0; S.rbp[j] = 0; S.rbd[j] = 0; S.type[j] = 0; }
        S.upre[j + 1] = S.upre[j] + (S.nrows[j] + S.rpu[j] - 1) / S.rpu[j];
    }
}
// unit u -> matrix j, first row (relative to this CTA's first row of j), row count
__device__ __forceinline__ void seg_unit(const SegTab & S, int u, int & j, int & off, int & n) {
    j = u < S.upre[1] ? 0 : u < S.upre[2] ? 1 : 2;
    off = (u - S.upre[j]) * S.rpu[j];
    n = min(S.rpu[j], S.nrows[j] - off);
}

#include "stream_dot.cuh"

// ---------------------------------------------------------------------------------------------------------------- block sync
// The CTA is warp-specialised: warps 0 .. SD_WARPS-1 are CONSUMERS (prologues, dot products, attention, grid barriers), warp SD_WARPS
// is the PRODUCER (phase staging + every bulk copy).  Consumers synchronise among themselves on named barrier 1.
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" :: "n"(SD_THREADS) : "memory"); }

__device__ __forceinline__ float cta_sum(float v, float * red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    cons_sync();
    // every thread adds the 12 warp sums itself, in warp order (broadcast LDS: no second shuffle tree, no trailing barrier — `red` is not
    // written again before the consumers have passed at least one more cons_sync)
    float t = red[0];
#pragma unroll
    for (int r = 1; r < SD_WARPS; ++r) t += red[r];
    return t;
}

// ---------------------------------------------------------------------------------------------------------------- prologues
// Build the q8 activation record of the phase's K-slice in shared memory (`act`).  The inputs were written by other CTAs in an earlier
// phase -> read through L2 (__ldcg).  x = x[0] + x[1] + ... in that fixed order (residual + K-split partials).
__device__ __forceinline__ float4 sd_load_x4(const SdPhase & P, int i) {
    float4 v = __ldcg((const float4 *) (P.x[0] + i));
    for (int s = 1; s < P.n_x; ++s) { const float4 a = __ldcg((const float4 *) (P.x[s] + i)); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
    return v;
}

// the shared-memory copy of the record keeps every plane 16-byte aligned (the global record packs d right after qs)
__device__ __forceinline__ ActLayout sd_act_layout(int act_group, int kl) {
    ActLayout L = act_layout(act_group == 256 ? B200_Q4_K : B200_Q8_0, kl);
    L.bsum_off = (L.bsum_off + 15) & ~(int64_t) 15;
    return L;
}

// fast path of sd_prologue (every phase of the decode program): ONE pass over the inputs — each warp keeps its <= 2 super-blocks in registers
// between the sum of squares and the quantisation.  A round trip to L2 costs ~1 us while the weight stream saturates HBM, so ALL loads of the
// prologue (NX summands x 2 blocks, and the norm weights) are issued back to back, branch-free, before any use.  NX is a template parameter:
// the single-input phases (wo, gate/up, down) must not carry the register pressure of the 4-summand layer-entry phase.
template <int NX>
__device__ __forceinline__ void sd_prologue_fast(const SdPhase & P, uint8_t * act, float * red, unsigned long long * pf, const ActLayout & L, int k, int kl, int k0,
                                                 bool norm, bool writer) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float scale = 1.0f;

    const int nb = kl >> 8, n_x = P.n_x;
    const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float4 a[2][NX][2], w[2][2];
    if (pf) pf[6] = globaltimer();
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const bool live = warp + SD_WARPS * t < nb;
        const int e = k0 + (live ? warp + SD_WARPS * t : warp) * 256 + lane * 8;
#pragma unroll
        for (int sx = 0; sx < NX; ++sx) {
            const bool on = live && sx < n_x;
            const float * px = P.x[on ? sx : 0] + e;
            a[t][sx][0] = z4; a[t][sx][1] = z4;
            if (on) ld256_cg(px, a[t][sx][0], a[t][sx][1]);              // ONE 256-bit request per summand and block: the prologue is bound by outstanding requests
        }
        w[t][0] = z4; w[t][1] = z4;
        if (norm && live) ld256_cg(P.norm_w + e, w[t][0], w[t][1]);
    }
    float v[2][8];
    float ss = 0.0f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        v[t][0] = a[t][0][0].x; v[t][1] = a[t][0][0].y; v[t][2] = a[t][0][0].z; v[t][3] = a[t][0][0].w;
        v[t][4] = a[t][0][1].x; v[t][5] = a[t][0][1].y; v[t][6] = a[t][0][1].z; v[t][7] = a[t][0][1].w;
#pragma unroll
        for (int sx = 1; sx < NX; ++sx) if (sx < n_x) {                 // fixed order x[0] + x[1] + ...: deterministic
            v[t][0] += a[t][sx][0].x; v[t][1] += a[t][sx][0].y; v[t][2] += a[t][sx][0].z; v[t][3] += a[t][sx][0].w;
            v[t][4] += a[t][sx][1].x; v[t][5] += a[t][sx][1].y; v[t][6] += a[t][sx][1].z; v[t][7] += a[t][sx][1].w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) ss += v[t][i] * v[t][i];
    }
    if (pf) pf[4] = globaltimer();
    if (norm) scale = 1.0f / sqrtf(cta_sum(ss, red) / (float) k + P.eps);
    if (pf) pf[5] = globaltimer();
const int blks[2] = { warp, warp + SD_WARPS };
    const bool lives[2] = { warp < nb, warp + SD_WARPS < nb };
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        if (lives[t]) {
            const int e = k0 + blks[t] * 256 + lane * 8;
            if (writer && P.x_out) { *(float4 *) (P.x_out + e) = make_float4(v[t][0], v[t][1], v[t][2], v[t][3]); *(float4 *) (P.x_out + e + 4) = make_float4(v[t][4], v[t][5], v[t][6], v[t][7]); }
            if (norm) {
                const float wv[8] = { w[t][0].x, w[t][0].y, w[t][0].z, w[t][0].w, w[t][1].x, w[t][1].y, w[t][1].z, w[t][1].w };
#pragma unroll
                for (int i = 0; i < 8; ++i) v[t][i] = __fmul_rn(__fmul_rn(v[t][i], scale), wv[i]);
                if (writer && P.norm_out) {
                    *(float4 *) (P.norm_out + e) = make_float4(v[t][0], v[t][1], v[t][2], v[t][3]); *(float4 *) (P.norm_out + e + 4) = make_float4(v[t][4], v[t][5], v[t][6], v[t][7]);
                }
            }
        }
    }
    quant_blocks_q8K<true, 2>(v, blks, lives, act, L.d_off, L.bsum_off);     // both blocks of the warp in lockstep (dummy second block for warps 4..11)
    cons_sync();
    }

__device__ __forceinline__ void sd_prologue(const SdPhase & P, int kpart, uint8_t * act, float * red, unsigned long long * pf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = P.k, kl = k / P.ksplit, k0 = kpart * kl;                 // this CTA quantises [k0, k0 + kl)
    const ActLayout L = sd_act_layout(P.act_group, kl);
    const bool writer = blockIdx.x == 0;
    if (P.prologue == SD_PRO_ACT) {
        const ActLayout G = act_layout(P.act_group == 256 ? B200_Q4_K : B200_Q8_0, kl);     // layout of the global record
        const uint4 * src = (const uint4 *) P.act;
        for (int i = threadIdx.x; i < (kl >> 4); i += SD_THREADS) *(uint4 *) (act + act_qs_off<true>(i * 16)) = __ldcg(src + i);
        const int dwords = (int) (G.bsum_off - G.d_off) >> 2, bwords = (int) (G.bytes - G.bsum_off) >> 2;    // small planes: 4-byte words
        for (int i = threadIdx.x; i < dwords; i += SD_THREADS) ((uint32_t *) (act + L.d_off))[i] = __ldcg((const uint32_t *) (P.act + G.d_off) + i);
        for (int i = threadIdx.x; i < bwords; i += SD_THREADS) ((uint32_t *) (act + L.bsum_off))[i] = __ldcg((const uint32_t *) (P.act + G.bsum_off) + i);
        cons_sync();
        return;
    }
    float scale = 1.0f;
    const bool norm = P.prologue == SD_PRO_RMSNORM_QUANT;
    if (P.act_group == 256 && (kl >> 8) <= 2 * SD_WARPS && (P.ksplit == 1 || !norm)) {
        if (P.n_x == 1) sd_prologue_fast<1>(P, act, red, pf, L, k, kl, k0, norm, writer);
        else            sd_prologue_fast<4>(P, act, red, pf, L, k, kl, k0, norm, writer);
        return;
    }
    if (norm) {                                                           // generic path: separate sum-of-squares pass over the FULL row
        float ss = 0.0f;
        for (int i = threadIdx.x * 4; i < k; i += SD_THREADS * 4) { const float4 v = sd_load_x4(P, i); ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
        scale = 1.0f / sqrtf(cta_sum(ss, red) / (float) k + P.eps);
    }
    if (P.act_group == 256) {
        for (int blk = warp; blk < (kl >> 8); blk += SD_WARPS) {
            const int e = k0 + blk * 256 + lane * 8;
            const float4 v0 = sd_load_x4(P, e), v1 = sd_load_x4(P, e + 4);
            float v[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
            if (writer && P.x_out) { *(float4 *) (P.x_out + e) = v0; *(float4 *) (P.x_out + e + 4) = v1; }
            if (norm) {
                const float4 w0 = __ldg((const float4 *) (P.norm_w + e)), w1 = __ldg((const float4 *) (P.norm_w + e + 4));
                const float w[8] = { w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w };
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __fmul_rn(__fmul_rn(v[i], scale), w[i]);
                if (writer && P.norm_out) {
                    *(float4 *) (P.norm_out + e) = make_float4(v[0], v[1], v[2], v[3]); *(float4 *) (P.norm_out + e + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
            quant_block_q8K<true>(v, act, blk, L.d_off, L.bsum_off);
        }
    } else {
        const int nblk = kl >> 5;
        for (int b0 = warp * 4; b0 < nblk; b0 += SD_WARPS * 4) {
            const int blk = b0 + (lane >> 3);
            const bool live = blk < nblk;
            float4 v = make_float4(0, 0, 0, 0);
            if (live) {
                const int e = k0 + blk * 32 + (lane & 7) * 4;
                v = sd_load_x4(P, e);
                if (writer && P.x_out) *(float4 *) (P.x_out + e) = v;
                if (norm) {
                    const float4 w = __ldg((const float4 *) (P.norm_w + e));
                    v.x = __fmul_rn(__fmul_rn(v.x, scale), w.x); v.y = __fmul_rn(__fmul_rn(v.y, scale), w.y);
                    v.z = __fmul_rn(__fmul_rn(v.z, scale), w.z); v.w = __fmul_rn(__fmul_rn(v.w, scale), w.w);
                    if (writer && P.norm_out) *(float4 *) (P.norm_out + e) = v;
                }
            }
            quant_block_q8_0<true>(v, live, act, blk, L.d_off, L.bsum_off);
        }
    }
    cons_sync();
}

// ---------------------------------------------------------------------------------------------------------------- grid barrier
// A monotonic arrival counter in global memory: every CTA adds 1 (release) and polls the SAME word (acquire) until it reaches the phase's
// target — one hop after the last arrival, instead of "last arriver sees the returned count, then flips a generation flag" (two).  The
// counter is never reset: bar[32] holds the value it had when the launch began (written by CTA 0 at the very end of the previous launch),
// so a captured CUDA graph can replay the kernel without host help.  `done` (shared) tells the producer warp that this CTA's consumers
// have left phase `phase_done - 1` (its staging slot may be reused).
__device__ __forceinline__ void sd_grid_barrier(unsigned * bar, unsigned target, volatile int * done, int phase_done, bool sync_grid) {
    cons_sync();
    if (threadIdx.x == 0) {
        *done = phase_done;
        if (sync_grid) {
            // the CTA's writes are ordered before thread 0's release by the bar.sync above (cumulativity); the acquire load orders the other
            // CTAs' writes before the bar.sync below
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(bar) : "memory");
            unsigned v;                                                    // (polling relaxed + one acquire fence measured slower)
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int) (v - target) < 0);
        }
    }
    if (sync_grid) cons_sync();
}

#include "stream_attn.cuh"

// ---------------------------------------------------------------------------------------------------------------- the kernel
__device__ __forceinline__ void sd_copy_phase(SdPhase * dst, const SdPhase * src, int lane, int nlanes) {
    static_assert(SD_WARPS % 4 == 0, "rings are split evenly over the 4 producer warps");
    static_assert(sizeof(SdPhase) % 16 == 0, "SdPhase must be a multiple of 16 bytes");
    for (int i = lane; i < (int) (sizeof(SdPhase) / 16); i += nlanes) ((uint4 *) dst)[i] = ((const uint4 *) src)[i];
}

// what a consumer warp needs to know about the unit sitting in one of its ring slots (written by the producer before the copy is issued,
// published by the slot's full-barrier): 32 bytes
struct alignas(16) SlotDesc { uint32_t type_n; uint32_t sub_p, sub_d; uint32_t pad_; float * y; const float * resid; };

struct StagedPhase { SdPhase P; SegTab S; };                           // ring of SD_NPH entries in shared memory, entry of phase p: p % SD_NPH

// producer lane (= consumer warp index): request unit u of the staged phase into the warp's ring slot
__device__ __forceinline__ void sd_issue(const SegTab & S, int u, uint32_t dst, uint32_t bar, SlotDesc * desc, uint64_t pol, bool dry) {
    int j, off, n; seg_unit(S, u, j, off, n);
    const uint32_t sp = S.sub_p[j], sd = S.sub_d[j];
    const int64_t rbp = S.rbp[j], rbd = S.rbd[j];
    desc->type_n = (uint32_t) S.type[j] | ((uint32_t) n << 8); desc->sub_p = sp; desc->sub_d = sd;
    desc->y = S.y[j] + off; desc->resid = S.resid[j] ? S.resid[j] + off : nullptr;
    if (dry) { mbar_arrive(bar); return; }                                // experiment: control path without any weight traffic
    mbar_expect_tx(bar, (uint32_t) n * (sp + sd) * S.nm);                 // release: orders the descriptor before the consumer's acquire
    const uint8_t * pay = S.pay[j] + off * rbp;
    const uint8_t * dpl = S.dpl[j] + off * rbd;
    if (S.ksplit == 1) {                                                   // full rows are back to back: one copy per plane
        bulk_g2s(dst, pay, (uint32_t) n * sp, bar, pol); dst += n * sp;
        if (sd) { bulk_g2s(dst, dpl, (uint32_t) n * sd, bar, pol); dst += n * sd; }
        if (S.nm == 2) {
            bulk_g2s(dst, S.pay2 + off * rbp, (uint32_t) n * sp, bar, pol); dst += n * sp;
            if (sd) bulk_g2s(dst, S.dpl2 + off * rbd, (uint32_t) n * sd, bar, pol);
        }
    } else {                                                               // K-slices of consecutive rows are rbp apart
        for (int i = 0; i < n; ++i) { bulk_g2s(dst, pay + i * rbp, sp, bar, pol); dst += sp; }
        if (sd) for (int i = 0; i < n; ++i) { bulk_g2s(dst, dpl + i * rbd, sd, bar, pol); dst += sd; }
    }
}

// L2 prefetch of this CTA's whole share of a staged matvec phase (all 32 lanes of the staging warp, fire-and-forget bulk prefetches): weights
// depend on nothing, so HBM can run `ahead` phases in front of the shared-memory rings — through the consumers' grid barriers, prologues and the
// attention phase, when the rings are full and the demand stream would otherwise stall.  The later demand copy (L2::evict_first) then hits L2.
__device__ __forceinline__ void l2_prefetch(const uint8_t * p, uint32_t bytes, uint64_t pol, bool hint) {
    if (hint) asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" :: "l"(p), "r"(bytes), "l"(pol) : "memory");
    else      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void l2_prefetch_span(const uint8_t * base, int64_t bytes, int lane, uint64_t pol, bool hint) {
    constexpr int CH = 4096;
    for (int64_t off = (int64_t) lane * CH; off < bytes; off += 32 * CH) l2_prefetch(base + off, (uint32_t) (bytes - off < CH ? bytes - off : CH), pol, hint);
}
__device__ __forceinline__ void sd_prefetch_l2(const SegTab & S, int lane, int shift, uint64_t pol, bool hint) {
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {
        const int nr = S.nrows[j] >> shift;                               // shift > 0: only the first 1/2^shift of the rows (byte budget)
        if (nr <= 0) continue;
        if (S.ksplit == 1) {
            l2_prefetch_span(S.pay[j], (int64_t) nr * S.rbp[j], lane, pol, hint);
            if (S.sub_d[j]) l2_prefetch_span(S.dpl[j], (int64_t) nr * S.rbd[j], lane, pol, hint);
            if (S.nm == 2) {
                l2_prefetch_span(S.pay2, (int64_t) nr * S.rbp[j], lane, pol, hint);
                if (S.sub_d[j]) l2_prefetch_span(S.dpl2, (int64_t) nr * S.rbd[j], lane, pol, hint);
            }
        } else {
            for (int r = lane; r < nr; r += 32) {
                l2_prefetch(S.pay[j] + (int64_t) r * S.rbp[j], (uint32_t) S.sub_p[j], pol, hint);
                if (S.sub_d[j]) l2_prefetch(S.dpl[j] + (int64_t) r * S.rbd[j], (uint32_t) S.sub_d[j], pol, hint);
            }
        }
    }
}

// The producer warp group (4 warps): warp 0 of the group stages phase descriptors + this CTA's segment tables two phases ahead of the one
// being issued; every producer warp keeps the rings of 3 consumer warps full (lanes 0..2, one ring each — bulk copies are issued from
// uniform registers, i.e. serially per lane, so the rings are spread over four warps rather than twelve lanes of one).  Producers never
// wait for a grid barrier (weights depend on nothing), only for free ring slots, so HBM keeps streaming through the consumers' barriers,
// prologues and attention phases.
__device__ __forceinline__ void sd_producer(const SdPhase * src, int n_phases, StagedPhase * ring_ph, volatile int * staged, volatile int * done,
                                            uint32_t ring, uint32_t full, uint32_t empty, SlotDesc * descs, uint64_t pol, uint32_t inflight, const SdRuntime & rt) {
    const int lane = threadIdx.x & 31, pw = (threadIdx.x >> 5) - SD_WARPS;
    constexpr int RPW = SD_WARPS / 4;                                      // rings per producer warp
    const int cw = pw * RPW + lane;                                        // the consumer warp this lane feeds (lanes < RPW)
    int staged_n = 0;
    uint32_t issued = 0;
    // experiment switches (profiling instantiation only; rt.flags == 0 in production): bit 10 = L2-prefetch every staged matvec phase,
    // bit 11 = with an evict_last hint, bits 13-14 = stage (and prefetch) this many phases ahead instead of SD_STAGE_AHEAD, bits 15-16 = only
    // the first 1/2, 1/4, 1/8 of each phase's rows
    const bool l2pf = (rt.flags & 1024) != 0, l2hint = (rt.flags & 2048) != 0;
    const int ahead = ((rt.flags >> 13) & 3) ? ((rt.flags >> 13) & 3) : SD_STAGE_AHEAD, l2shift = (rt.flags >> 15) & 3;
    uint64_t pol_last = 0;
    if (l2hint) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    for (int p = 0; p < n_phases; ++p) {
        if (pw == 0) {
            while (staged_n < n_phases && staged_n <= p + ahead) {
                if (staged_n >= SD_NPH) while (*done < staged_n - SD_NPH + 1) { }      // the slot's previous tenant has been left by the consumers
                StagedPhase & E = ring_ph[staged_n % SD_NPH];
                sd_copy_phase(&E.P, src + staged_n, lane, 32);
                __syncwarp();
                // an attention phase's old KV rows do not depend on this token: pull this CTA's chunk into L2 now, two phases early
                if (E.P.kind == SD_ATTN && !(rt.flags & 2)) sa_prefetch_kv(E.P.attn, rt, lane, 32);
                if (lane == 0) {
                    if (E.P.kind == SD_MATVEC) seg_build(E.S, E.P, blockIdx.x, gridDim.x); else E.S.upre[3] = 0;
                    __threadfence_block();
                    *staged = staged_n + 1;
                }
                __syncwarp();
                if (l2pf && E.P.kind == SD_MATVEC) sd_prefetch_l2(E.S, lane, l2shift, pol_last, l2hint);
                ++staged_n;
            }
        } else {
            while (*staged <= p) { }
            __threadfence_block();
        }
        const StagedPhase & E = ring_ph[p % SD_NPH];
        if (E.P.kind != SD_MATVEC) continue;
        const int nunits = E.S.upre[3];
        if (lane < RPW) {
            const uint32_t ring_w = ring + cw * SD_DEPTH * SD_SLOT_BYTES, full_w = full + cw * SD_DEPTH * 8, empty_w = empty + cw * SD_DEPTH * 8;
            for (int u = cw; u < nunits; u += SD_WARPS) {
                const uint32_t slot = issued % SD_DEPTH, use = issued / SD_DEPTH;
                if (use) mbar_wait(empty_w + 8 * slot, (use - 1) & 1);
                // pacing: at most `inflight` units of this ring on the wire
                if (issued >= inflight) { const uint32_t o = issued - inflight; mbar_wait(full_w + 8 * (o % SD_DEPTH), (o / SD_DEPTH) & 1); }
                sd_issue(E.S, u, ring_w + slot * SD_SLOT_BYTES, full_w + 8 * slot, descs + cw * SD_DEPTH + slot, pol, (rt.flags & 4) != 0);
                ++issued;
            }
        }
        __syncwarp();
    }
}

// one matvec phase on a consumer warp.  REGS: K-slice <= 4096 and q4_K / q6_K only -> activation fragments live in registers; otherwise
// fragments are re-read from shared memory (two rows share each read).
template <bool REGS>
__device__ __forceinline__ void sd_consume(const SdPhase & P, int nunits, const ActS & A, uint32_t & consumed, uint32_t ring_w, uint32_t full_w, uint32_t empty_w,
                                           const SlotDesc * descs_w, unsigned long long * pf, int flags) {
    long long waited = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kl = P.k / P.ksplit, nblk = kl >> 8;
    const bool swiglu = P.epilogue == SD_EPI_SWIGLU;
    HFrag fr;
    if (REGS) hfrag_fill(A, nblk, fr);
    if (pf && (flags & 512)) pf[4] = globaltimer();                        // finer marks of warp 0's first unit (B200_SD_FLAGS=512)
    bool first = true;
    for (int u = warp; u < nunits; u += SD_WARPS) {
        const uint32_t slot = consumed % SD_DEPTH, par = (consumed / SD_DEPTH) & 1;
        const uint32_t base = ring_w + slot * SD_SLOT_BYTES;
        const long long tw = pf ? clock64() : 0;
        mbar_wait(full_w + 8 * slot, par);
        if (pf) waited += clock64() - tw;
        if (pf && (flags & 512) && first) pf[5] = globaltimer();
        const SlotDesc d = descs_w[slot];
        const int type = d.type_n & 0xff, n = d.type_n >> 8;
        float rz = 0.0f;                                                   // the residual's L2 round trip overlaps the dot products
        if (d.resid && lane < n) rz = __ldcg(d.resid + lane);
        const uint32_t rbp = d.sub_p, rbd = d.sub_d;
        const uint32_t bp = (uint32_t) n * rbp, bd = (uint32_t) n * rbd;     // slot: [payload rows][d rows] (+ the same again for `up`)
        float2 g2 = make_float2(0.0f, 0.0f), v = make_float2(0.0f, 0.0f);
        if (flags & 8) { }                                                 // experiment: handshake only, no arithmetic
        else if (REGS) {
            if (type == B200_Q4_K) { v = krow_regs<B200_Q4_K>(base, rbp, base + bp, rbd, nblk, n, fr); if (swiglu) { g2 = v; v = krow_regs<B200_Q4_K>(base + bp + bd, rbp, base + 2 * bp + bd, rbd, nblk, n, fr); } }
            else                   { v = krow_regs<B200_Q6_K>(base, rbp, base + bp, rbd, nblk, n, fr); if (swiglu) { g2 = v; v = krow_regs<B200_Q6_K>(base + bp + bd, rbp, base + 2 * bp + bd, rbd, nblk, n, fr); } }
        } else {
            v = unit_dots_lds(type, n, base, rbp, base + bp, rbd, A, kl);
            if (swiglu) { g2 = v; v = unit_dots_lds(type, n, base + bp + bd, rbp, base + 2 * bp + bd, rbd, A, kl); }
        }
        if (pf && (flags & 512) && first) { pf[6] = globaltimer(); first = false; }
        __syncwarp();                                                      // every lane's reads of the slot are done
        if (lane == 0) mbar_arrive(empty_w + 8 * slot);                    // hand it back to the producer before the epilogue's global traffic
        if (lane < n) {
            float o = lane == 0 ? v.x : v.y;
            if (swiglu) { const float gg = lane == 0 ? g2.x : g2.y; o = (gg / (1.0f + expf(-gg))) * o; }
            o += rz;
            d.y[lane] = o;
        }
        ++consumed;
    }
    if (pf) pf[7] = (unsigned long long) waited;                           // cycles warp 0 spent waiting for weights in this phase
}

// PROF = false is the production instantiation: `pf` is a compile-time null in every inlined helper and the experiment switches are a compile-time
// 0, so neither the %globaltimer stamps nor their branches and registers exist in it (the kernel is far larger than the instruction caches
// and its speed moves by several per cent with the position of the hot loops).
template <bool PROF>
__global__ void __launch_bounds__(SD_THREADS + 128, 1) k_stream(const SdPhase * __restrict__ phases_g, int n_phases, unsigned * gbar,
                                                               const __grid_constant__ SdPhase single, const SdRuntime rt_in) {
    SdRuntime rt = rt_in;
    if (!PROF) { rt.flags = 0; rt.prof = nullptr; }
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t  * ring = smem;
    uint8_t  * act  = smem + SD_RING_BYTES;
    uint8_t  * attn_scratch = act + SD_ACT_BYTES;
    uint64_t * bars = (uint64_t *) (attn_scratch + SD_ATTN_BYTES);       // full[warp][slot], then empty[warp][slot]
    SlotDesc * descs = (SlotDesc *) (bars + 2 * SD_WARPS * SD_DEPTH);
    float    * red  = (float *) (descs + SD_WARPS * SD_DEPTH);
    StagedPhase * ring_ph = (StagedPhase *) (red + 64);
    float2   * rope_tab = (float2 *) (((uintptr_t) (ring_ph + SD_NPH) + 15) & ~(uintptr_t) 15);   // [64] (cos, sin) of the token's position
    volatile int * staged = (volatile int *) (rope_tab + 64);          // number of phases staged so far (producer -> consumers)
    volatile int * done   = staged + 1;                                // number of phases the consumers have left (consumers -> producer)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SdPhase * src = phases_g ? phases_g : &single;
    const uint32_t full = smem_u32(bars), empty = full + SD_WARPS * SD_DEPTH * 8;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * SD_WARPS * SD_DEPTH; ++s) mbar_init(full + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *staged = 0; *done = 0;
    }
    __syncthreads();                                                     // the only CTA-wide barrier: roles split here

    // register reallocation between the warp groups (512 threads start with 128 registers each): the producer group gives most of its
    // registers back, the three consumer groups grow to 152
    if (warp >= SD_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const uint64_t pol = rt.flags & 1 ? policy_evict_normal() : policy_evict_first();
        const uint32_t infl = (rt.flags >> 4) & 15;
        sd_producer(src, n_phases, ring_ph, staged, done, smem_u32(ring), full, empty, descs, pol, infl ? infl : SD_INFLIGHT, rt);
        return;
    }

    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    if (warp == 2 && rt.has_rope) sa_rope_table(rope_tab, rt, lane);     // published by the first prologue's barrier
    const uint32_t ring_w = smem_u32(ring) + warp * SD_DEPTH * SD_SLOT_BYTES, full_w = full + warp * SD_DEPTH * 8, empty_w = empty + warp * SD_DEPTH * 8;
    const SlotDesc * descs_w = descs + warp * SD_DEPTH;
    uint32_t consumed = 0;
    unsigned bar_base = 0;                                                // value of the arrival counter when this launch began
    if (threadIdx.x == 0 && n_phases > 1) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(bar_base) : "l"(gbar + 32) : "memory");
    const bool prof = PROF && rt.prof != nullptr && threadIdx.x == 0;
    unsigned long long * const pbase = prof ? rt.prof + (size_t) blockIdx.x * 8 : nullptr;      // [phase][cta][8]
    const size_t pstr = (size_t) gridDim.x * 8;

    for (int p = 0; p < n_phases; ++p) {
        if (prof) pbase[p * pstr + 0] = globaltimer();
        while (*staged <= p) { }
        __threadfence_block();
        const StagedPhase & E = ring_ph[p % SD_NPH];
        const SdPhase & P = E.P;
        if (P.kind == SD_MATVEC) {
            sd_prologue(P, E.S.kpart, act, red, PROF && prof ? pbase + p * pstr : nullptr);
            if (prof) pbase[p * pstr + 1] = globaltimer();
            const int kl = P.k / P.ksplit;
            const ActLayout L = sd_act_layout(P.act_group, kl);
            ActS A; A.qs = smem_u32(act); A.d = A.qs + (uint32_t) L.d_off; A.bsum = A.qs + (uint32_t) L.bsum_off;
            bool regs = P.act_group == 256 && kl <= 4096;                 // q8_K fragments of a 4096-wide record fit in registers
            for (int m = 0; m < P.n_mat; ++m) regs = regs && (P.mat[m].type == B200_Q4_K || P.mat[m].type == B200_Q6_K);
            if (regs) sd_consume<true >(P, E.S.upre[3], A, consumed, ring_w, full_w, empty_w, descs_w, PROF && prof ? pbase + p * pstr : nullptr, rt.flags);
            else      sd_consume<false>(P, E.S.upre[3], A, consumed, ring_w, full_w, empty_w, descs_w, PROF && prof ? pbase + p * pstr : nullptr, rt.flags);
        } else {
            if (prof) pbase[p * pstr + 1] = globaltimer();
            sd_attention(P, rt, attn_scratch, rope_tab, PROF && prof ? pbase + p * pstr : nullptr);
        }
        if (prof) pbase[p * pstr + 2] = globaltimer();
        sd_grid_barrier(gbar, bar_base + (unsigned) (p + 1) * gridDim.x, done, p + 1, p + 1 < n_phases);
        if (prof) pbase[p * pstr + 3] = globaltimer();
    }
    // every CTA read bar[32] before its first barrier, and CTA 0 is past the last one: safe to publish the next launch's base
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_phases > 1) gbar[32] = bar_base + (unsigned) (n_phases - 1) * gridDim.x;
}

// ---------------------------------------------------------------------------------------------------------------- host side
bool sd_fill_mat(SdMat & M, const void * w, int type, int layout, int64_t m, int64_t k, int64_t row_stride_bytes, float * y, const float * residual) {
    if (!is_quant(type) || k % blck_size(type) || m <= 0 || m > INT32_MAX) return false;
    const int64_t nblk = k / blck_size(type);
    const bool split = payload_size(type) != type_size(type);            // q4_0 / q8_0 / q6_K: streamable only from the planar layout
    if (split && layout != B200_LAYOUT_PLANAR) return false;
    const int64_t rb_p = nblk * payload_size(type), rb_d = split ? nblk * 2 : 0;
    if (!split && row_stride_bytes != rb_p) return false;                 // rows must be back to back for 1-D bulk copies
    if (rb_p % 16 || rb_d % 16 || (uintptr_t) w % 16) return false;
    M.payload = (const uint8_t *) w; M.dplane = split ? (const uint8_t *) w + m * rb_p : nullptr;
    if (split && ((uintptr_t) M.dplane % 16)) return false;
    M.y = y; M.residual = residual; M.type = type; M.rows = (int32_t) m; M.row_bytes_p = (int32_t) rb_p; M.row_bytes_d = (int32_t) rb_d;
    return true;
}

// a phase is streamable when one row's K-slice (x2 for gate/up pairs) fits a ring slot and every slice is 16-byte granular
bool sd_phase_ok(const SdPhase & P) {
    if (P.kind != SD_MATVEC) return true;
    const int ks = P.ksplit;
    if (ks < 1 || P.k % ks) return false;
    const int64_t kl = P.k / ks;
    if (act_layout(P.act_group == 256 ? B200_Q4_K : B200_Q8_0, kl).bytes + 16 > SD_ACT_BYTES) return false;
    const int nm = P.epilogue == SD_EPI_SWIGLU ? 2 : 1;
    for (int j = 0; j < P.n_mat; ++j) {
        const SdMat & M = P.mat[j];
        if (M.row_bytes_p % ks || M.row_bytes_d % ks) return false;
        const int sp = M.row_bytes_p / ks, sd = M.row_bytes_d / ks;
        if (sp % 16 || sd % 16 || (sp + sd) * nm > SD_SLOT_BYTES) return false;
        if (ks > 1 && kl % 256) return false;                              // K-split: whole super-blocks (8 q4_0 / q8_0 blocks) per slice
        if ((is_kquant(M.type) ? 256 : 32) != P.act_group) return false;
    }
    if (nm == 2 && (P.n_mat != 2 || P.mat[0].type != P.mat[1].type || P.mat[0].rows != P.mat[1].rows || P.mat[0].row_bytes_p != P.mat[1].row_bytes_p)) return false;
    return true;
}

static int sd_setup(bool prof) {
    static smem_mask_t done[2] = { {0}, {0} };
    const cudaError_t e = prof ? ensure_dyn_smem(k_stream<true>, SD_SMEM_BYTES, done[1]) : ensure_dyn_smem(k_stream<false>, SD_SMEM_BYTES, done[0]);
    return e == cudaSuccess ? B200_OK : -(int) e;
}

// one launch of the persistent kernel.  phases_dev == nullptr: run `single` (no grid barrier needed -> ordinary launch);
// otherwise a cooperative launch guarantees the one-CTA-per-SM grid is co-resident for the grid barriers.
int sd_launch(const SdPhase * phases_dev, int n_phases, const SdPhase * single, unsigned * gbar, const SdRuntime & rt, cudaStream_t st) {
    static const int env_flags = getenv("B200_SD_FLAGS") ? atoi(getenv("B200_SD_FLAGS")) : 0;     // experiment switches (bit 0: no evict_first hint)
    const bool prof = rt.prof != nullptr || env_flags != 0;
    int rc = sd_setup(prof);
    if (rc) return rc;
    static const SdPhase zero = {};
    const SdPhase & S = single ? *single : zero;
    SdRuntime rt2 = rt; rt2.flags = env_flags;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned) sm_count()); cfg.blockDim = dim3(SD_THREADS + 128); cfg.dynamicSmemBytes = SD_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = (phases_dev && n_phases > 1) ? 1 : 0;
    if (prof) B200_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_stream<true>, phases_dev, n_phases, gbar, S, rt2));
    else      B200_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_stream<false>, phases_dev, n_phases, gbar, S, rt2));
    return B200_OK;
}

} // namespace b200
