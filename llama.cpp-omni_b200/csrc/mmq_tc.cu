// mmq_tc.cu — prefill / batched GEMM  Y[m, n] = W[m, k] . X[n, k]^T  for quantised W (q4_K / q5_K native, q6_K / q8_0 / q4_0 planar) and n > 8
// columns,
// on the 5th-generation tensor cores: tcgen05.mma (kind::f16, M = 128, N <= 256, K = 16) with the accumulator in TMEM.
//
// Replaces ggml_cuda_mul_mat_q (ggml-cuda/mmq.cu:205, kernel mul_mat_q mmq.cuh:3136 with the q4_K / q6_K tile loaders :1741, :2043, and
// quantize_mmq_q8_1 quantize.cu:50-146) — int8 mma.sync tiles in the reference; here the quant blocks are dequantised to F16 in shared
// memory and the contraction runs as F16 x F16 -> F32 UMMA.  K-quant sub-block scales (6-bit scale/min per 32 weights, int8 scale per 16)
// do not map onto the hardware block-scaled formats, so the dequant has to materialise F16 (SURVEY.md §7 "tcgen05 prefill").
//
// One persistent CTA per SM, 15 warps, warp-specialised (the canonical Blackwell GEMM anatomy, hand-written in PTX):
//   warps 0-7   A producers: copy their half of a raw quant block from the raw ring to registers (thread group g = warps 4g..4g+3 owns half g of
//               every 256-weight block = K-chunks 2g, 2g+1, so the groups fill alternate pairs of stages), dequantise with half2 math
//               (byte -> 1024+q via PRMT with 0x64, HSUB2, HFMA2 by d*sc / -dmin*m) and store the 128 x 64 F16 tile in the canonical K-major
//               no-swizzle UMMA layout (8 x 16-byte core matrices; row groups 128 B apart -> conflict-free 16-byte stores);
//               fence.proxy.async + mbarrier arrive hands the stage to the tensor core
//   warp  8     B producer: the activations were converted F32 -> F16 ONCE by k_x_to_f16_tiles into the same canonical layout, tile by tile,
//               so a 256 x 64 B tile is plain contiguous memory: 4 x 8 KB cp.async.bulk (TMA) per stage, complete_tx on the stage barrier
//   warp  9     MMA issuer: one lane issues 4 tcgen05.mma per stage (D in TMEM, 2 x 256 columns double-buffered) and tcgen05.commit's the stage
//               back to the producers / the accumulator to the epilogue
//   warp  14    raw weight producer: ONE 2-D tiled TMA (cp.async.bulk.tensor, box = 128 rows x one 144 / 208-byte quant block) per 4 stages
//               into a 3-slot raw ring, up to 3 blocks ahead of the dequantisers -> no register scoreboard ever waits on L2 / HBM
//   warps 10-13 epilogue: tcgen05.ld 32 lanes x 32 columns, coalesced F32 stores (lane = weight row = contiguous dst index)
// FLOPs per launch = 2 m n k.  Tensor-bound for n >= ~64.
//
// Numerics: weights are rounded to F16 after dequantisation (relative 2^-11 per factor), activations to F16, products accumulate in F32.
// The CPU oracle instead quantises activations to q8_K (~3e-3 relative per element): this path is closer to the exact product than the
// oracle is; parity bar = the reference's own MUL_MAT bar, NMSE <= 5e-4 (tests/test-backend-ops.cpp:3300).
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace b200 {

constexpr int TC_M = 128, TC_N = 256, TC_K = 64, TC_STAGES = 3;
constexpr int TC_RAW_REGION = 3 * TC_M * 208;                                               // raw quant-block ring: 128 rows x the bytes of 256 weights per slot
constexpr int TC_RAW_MAX = 3;                                                               // slots: 3 (2 for q8_0, whose 256-weight span is 256 B per row)
constexpr int TC_A_BYTES = TC_M * TC_K * 2, TC_B_BYTES = TC_N * TC_K * 2;                 // 16 KB, 32 KB
constexpr int TC_A_LBO = (TC_M / 8) * 128, TC_B_LBO = (TC_N / 8) * 128, TC_SBO = 128;        // core matrices: K direction / row-group direction
constexpr int TC_DEQ_THREADS = 256, TC_THREADS = 480;
constexpr int TC_SMEM = TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + TC_RAW_REGION + 256;
static_assert(TC_SMEM <= 227 * 1024, "k_mmq_tc shared memory");

struct TcArgs {
    const uint8_t * w; const uint8_t * wd;          // payload plane, f16 d plane (q6_K planar) or null
    const uint8_t * x16;                            // activations, F16, pre-tiled: [n_tiles][k/64][TC_B_BYTES]
    float * dst; int64_t dst_ld;                    // dst[n * dst_ld + m]
    const float * resid;                            // optional: dst = W.x + resid (same leading dimension as dst; split-K launches never carry it)
    int64_t m, k, n, row_bytes;                     // n = real columns; row_bytes of the payload plane
    int type, tiles_m, tiles_n;
    int span_bytes, n_raw;                          // payload bytes of 256 weights of one row (= TMA box width); raw ring slots
    int splitk;                                     // 1, or 2: two CTAs share a tile, each reduces half of K and adds its partial to dst with red.global.add (dst zeroed by the launcher;
                                                    // two addends commute, so the result is still deterministic)
    int nsub;                                       // 1: tiles are 256 columns wide; 2: 128 (twice the tiles when 256-wide ones cannot fill the SMs)
    int flags;                                      // experiment switches (B200_TC_FLAGS): 1 no dequant, 2 no MMA, 4 no B copies
    // several weight matrices that multiply the SAME activations in one launch (q / k / v: 48 m-tiles instead of 32 + 8 + 8, so the 1024-row wk / wv no longer run
    // as two launches that cannot fill the SMs).  Segment s owns the m-tiles [seg[s].tile0, seg[s + 1].tile0); the fields above describe segment 0 / the totals.
    int nseg, raw_stride;                           // raw_stride: bytes of one raw-ring slot (128 rows x the LARGEST 256-weight span of the launch)
    struct Seg { const uint8_t * wd; float * dst; int64_t dst_ld, m; int type, span_bytes, tile0, pad_; } seg[3];
};
// the segment an m-tile belongs to
__device__ __forceinline__ int tc_seg_of(const TcArgs & A, int mt) { return A.nseg > 2 && mt >= A.seg[2].tile0 ? 2 : A.nseg > 1 && mt >= A.seg[1].tile0 ? 1 : 0; }

// ---------------------------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t tc_smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void tc_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_bulk_g2s(uint32_t dst, const void * src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// 2-D tiled TMA: box {block bytes, 128 rows} of the weight payload plane -> dense [128][block bytes] in shared memory (rows past m read as 0)
__device__ __forceinline__ void tc_tma_2d(uint32_t dst, const CUtensorMap * map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ uint4 tc_lds16(uint32_t a) { uint4 r; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a)); return r; }
__device__ __forceinline__ uint2 tc_lds8(uint32_t a) { uint2 r; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a)); return r; }
// one stage of activations: the 256-column tile is [8 K-chunks][32 row groups][128 B]; a 128-column half tile is 8 pieces of 2 KB
__device__ __forceinline__ void tc_load_b(uint32_t dst, const uint8_t * tile, int nsub, int half, uint32_t bar, bool expect) {
    if (nsub == 1) {
        if (expect) tc_mbar_expect_tx(bar, 32768);
#pragma unroll
        for (int q = 0; q < 4; ++q) tc_bulk_g2s(dst + q * 8192, tile + q * 8192, 8192, bar);
    } else {
        if (expect) tc_mbar_expect_tx(bar, 16384);
#pragma unroll
        for (int q = 0; q < 8; ++q) tc_bulk_g2s(dst + q * 2048, tile + q * 4096 + half * 2048, 2048, bar);
    }
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, cute/arch/mma_sm100_desc.hpp): start address, leading
// (K direction) and stride (M/N direction) byte offsets in 16-byte units, descriptor version 1 (Blackwell), layout type 0
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t) ((saddr & 0x3ffff) >> 4) | ((uint64_t) (lbo >> 4) << 16) | ((uint64_t) (sbo >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t tc_idesc(int n) { return (1u << 4) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (TC_M >> 4) << 24); }
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                 "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------- dequant -> F16
// 4 bytes (values 0..255 each) -> two half2 holding 1024 + byte: PRMT interleaves the bytes with 0x64 (f16 0x64xx = 1024 + xx)
__device__ __forceinline__ void bytes_to_h2(uint32_t b, __half2 & lo, __half2 & hi) {
    const uint32_t x = __byte_perm(b, 0x64646464u, 0x4140), y = __byte_perm(b, 0x64646464u, 0x4342);
    lo = *(const __half2 *) &x; hi = *(const __half2 *) &y;
}
__device__ __forceinline__ uint4 ldg16(const uint8_t * p) { return __ldg((const uint4 *) p); }

// 8 consecutive weights w = (q - off) * s + c from two words of byte codes -> one 16-byte core-matrix row
__device__ __forceinline__ uint4 deq8(uint32_t b0, uint32_t b1, __half2 off, __half2 s, __half2 c) {
    __half2 h[4];
    bytes_to_h2(b0, h[0], h[1]); bytes_to_h2(b1, h[2], h[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __hfma2(__hsub2(h[i], off), s, c);
    return *(const uint4 *) h;
}

// Producer mapping: thread (row, g) owns HALF g of every 256-weight block of its row = K-chunks 2g and 2g+1 (64 weights each), so the two
// thread groups fill alternate PAIRS of stages and no byte is loaded twice.  The raw bytes of a half block (80 B q4_K, 116 B q6_K) are
// loaded into registers TWO blocks (8 stages, ~3.5 us of MMA time) ahead of their use: the L2/HBM round trip never stalls the pipeline.
struct RawQ4K { uint4 hdr, q[4]; };                       // d|dmin + 12 scale bytes; qs[64g .. 64g+64)
struct RawQ6K { uint4 l[4], h[2]; uint2 sc; float d; };   // ql[64g ..+64), qh[32g ..+32), int8 scales[8g ..+8), d

// blk = shared address of this row's block in the raw ring slot (row stride 144 / 208 B: 16-byte reads of 8 consecutive rows hit 8 distinct
// 16-byte bank groups -> conflict-free)
__device__ __forceinline__ void raw_load(RawQ4K & R, uint32_t blk, int g) {
    R.hdr = tc_lds16(blk);
#pragma unroll
    for (int i = 0; i < 4; ++i) R.q[i] = tc_lds16(blk + 16 + 64 * g + 16 * i);
}
__device__ __forceinline__ void raw_load(RawQ6K & R, uint32_t pay, int g) {
#pragma unroll
    for (int i = 0; i < 4; ++i) R.l[i] = tc_lds16(pay + 64 * g + 16 * i);
    R.h[0] = tc_lds16(pay + 128 + 32 * g); R.h[1] = tc_lds16(pay + 128 + 32 * g + 16);
    R.sc = tc_lds8(pay + 192 + 8 * g);
}

struct RawQ5K { uint4 hdr, h[2], q[4]; };                 // d|dmin + scales; qh[32]; qs[64g ..+64)
struct RawQ80 { uint4 q[8]; float d[4]; };                // int8 of blocks 4g .. 4g+3 (planar payload); their f16 d (d plane)
struct RawQ40 { uint4 q[4]; float d[4]; };                // nibbles of blocks 4g .. 4g+3; their d
__device__ __forceinline__ void raw_load(RawQ5K & R, uint32_t blk, int g) {
    R.hdr = tc_lds16(blk); R.h[0] = tc_lds16(blk + 16); R.h[1] = tc_lds16(blk + 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) R.q[i] = tc_lds16(blk + 48 + 64 * g + 16 * i);
}
__device__ __forceinline__ void raw_load(RawQ80 & R, uint32_t pay, int g) {
#pragma unroll
    for (int i = 0; i < 8; ++i) R.q[i] = tc_lds16(pay + 128 * g + 16 * i);
}
__device__ __forceinline__ void raw_load(RawQ40 & R, uint32_t pay, int g) {
#pragma unroll
    for (int i = 0; i < 4; ++i) R.q[i] = tc_lds16(pay + 64 * g + 16 * i);
}

__device__ __forceinline__ void st_row(uint32_t addr, const uint4 & v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// thread (row, g) writes the 64 weights of its cc-th chunk (block chunk c = 2g + cc) into stage memory (a_row = shared address of its row)
template <int cc> __device__ __forceinline__ void deq_chunk(const RawQ4K & R, int g, uint32_t a_row) {
    // chunk c: low nibbles of qs[32c ..+32) = sub-block 2c (k 0..31 of the chunk), high nibbles = sub-block 2c+1 (k 32..63);
    // their 6-bit scales / mins (get_scale_min_k4, ggml-quants.c:703-711) are bytes 2cc, 2cc+1 of the group's packed word
    const uint4 & hdr = R.hdr;
    const uint32_t scw = g ? ((hdr.w & 0x0f0f0f0fu) | (((hdr.y >> 6) & 0x03030303u) << 4)) : (hdr.y & 0x3f3f3f3fu);
    const uint32_t mnw = g ? (((hdr.w >> 4) & 0x0f0f0f0fu) | (((hdr.z >> 6) & 0x03030303u) << 4)) : (hdr.z & 0x3f3f3f3fu);
    const float d = h2f(hdr.x & 0xffff), dmin = h2f(hdr.x >> 16);
    const __half2 off = __float2half2_rn(1024.0f);
    const uint32_t w[8] = { R.q[2 * cc].x, R.q[2 * cc].y, R.q[2 * cc].z, R.q[2 * cc].w, R.q[2 * cc + 1].x, R.q[2 * cc + 1].y, R.q[2 * cc + 1].z, R.q[2 * cc + 1].w };
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        const float sc = (float) ((scw >> (16 * cc + 8 * hf)) & 0xff), mn = (float) ((mnw >> (16 * cc + 8 * hf)) & 0xff);
        const __half2 s = __float2half2_rn(d * sc), cm = __float2half2_rn(-(dmin * mn));
#pragma unroll
        for (int i = 0; i < 4; ++i)
            st_row(a_row + (4 * hf + i) * TC_A_LBO, deq8((w[2 * i] >> (4 * hf)) & 0x0f0f0f0fu, (w[2 * i + 1] >> (4 * hf)) & 0x0f0f0f0fu, off, s, cm));
    }
}

template <int cc> __device__ __forceinline__ void deq_chunk(const RawQ6K & R, int g, uint32_t a_row) {
    // half g of the block, chunk cc = quads 2cc and 2cc+1 (dequantize_row_q6_K, ggml-quants.c:1762-1791): quad q reads ql[64g + 32(q&1) + l]
    // (low nibble for quads 0/1, high for 2/3) and bits 2q, 2q+1 of qh[32g + l]; int8 scale index 8g + 2q + (l >> 4)
    const uint32_t hw[8] = { R.h[0].x, R.h[0].y, R.h[0].z, R.h[0].w, R.h[1].x, R.h[1].y, R.h[1].z, R.h[1].w };
    const uint32_t scw = cc ? R.sc.y : R.sc.x;            // scales[8g + 4cc .. +3]: quad 2cc -> bytes 0, 1; quad 2cc+1 -> bytes 2, 3
    const __half2 off = __float2half2_rn(1056.0f), zero = __float2half2_rn(0.0f);        // 1024 (PRMT bias) + 32 (q6_K code offset)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {                     // quad = 2cc + hf -> k 32hf .. 32hf+31 of the chunk; ql bytes 32hf ..+32 of the half
        const float sc0 = (float) (int) (int8_t) (scw >> (16 * hf)), sc1 = (float) (int) (int8_t) (scw >> (16 * hf + 8));
        const __half2 s0 = __float2half2_rn(R.d * sc0), s1 = __float2half2_rn(R.d * sc1);
        const uint32_t lw[8] = { R.l[2 * hf].x, R.l[2 * hf].y, R.l[2 * hf].z, R.l[2 * hf].w, R.l[2 * hf + 1].x, R.l[2 * hf + 1].y, R.l[2 * hf + 1].z, R.l[2 * hf + 1].w };
        constexpr int shl = 4 * cc;
        const int shh = 2 * (2 * cc + hf);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t b0 = ((lw[2 * i] >> shl) & 0x0f0f0f0fu) | (((hw[2 * i] >> shh) & 0x03030303u) << 4);
            const uint32_t b1 = ((lw[2 * i + 1] >> shl) & 0x0f0f0f0fu) | (((hw[2 * i + 1] >> shh) & 0x03030303u) << 4);
            st_row(a_row + (4 * hf + i) * TC_A_LBO, deq8(b0, b1, off, i < 2 ? s0 : s1, zero));
        }
    }
}

template <int cc> __device__ __forceinline__ void deq_chunk(const RawQ5K & R, int g, uint32_t a_row) {
    // q4_K plus one high bit per weight: group j = 2g + cc uses bit 2j of qh[l] for its low-nibble sub-block and bit 2j+1 for the high one
    // (dequantize_row_q5_K, ggml-quants.c:1554-1578)
    const uint4 & hdr = R.hdr;
    const uint32_t scw = g ? ((hdr.w & 0x0f0f0f0fu) | (((hdr.y >> 6) & 0x03030303u) << 4)) : (hdr.y & 0x3f3f3f3fu);
    const uint32_t mnw = g ? (((hdr.w >> 4) & 0x0f0f0f0fu) | (((hdr.z >> 6) & 0x03030303u) << 4)) : (hdr.z & 0x3f3f3f3fu);
    const float d = h2f(hdr.x & 0xffff), dmin = h2f(hdr.x >> 16);
    const __half2 off = __float2half2_rn(1024.0f);
    const uint32_t w[8] = { R.q[2 * cc].x, R.q[2 * cc].y, R.q[2 * cc].z, R.q[2 * cc].w, R.q[2 * cc + 1].x, R.q[2 * cc + 1].y, R.q[2 * cc + 1].z, R.q[2 * cc + 1].w };
    const uint32_t hw[8] = { R.h[0].x, R.h[0].y, R.h[0].z, R.h[0].w, R.h[1].x, R.h[1].y, R.h[1].z, R.h[1].w };
    const int j = 2 * g + cc;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        const float sc = (float) ((scw >> (16 * cc + 8 * hf)) & 0xff), mn = (float) ((mnw >> (16 * cc + 8 * hf)) & 0xff);
        const __half2 s = __float2half2_rn(d * sc), cm = __float2half2_rn(-(dmin * mn));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t b0 = ((w[2 * i] >> (4 * hf)) & 0x0f0f0f0fu) | (((hw[2 * i] >> (2 * j + hf)) & 0x01010101u) << 4);
            const uint32_t b1 = ((w[2 * i + 1] >> (4 * hf)) & 0x0f0f0f0fu) | (((hw[2 * i + 1] >> (2 * j + hf)) & 0x01010101u) << 4);
            st_row(a_row + (4 * hf + i) * TC_A_LBO, deq8(b0, b1, off, s, cm));
        }
    }
}

template <int cc> __device__ __forceinline__ void deq_chunk(const RawQ80 & R, int, uint32_t a_row) {
    // blocks 2cc, 2cc+1 of the thread's four: w = d * int8 (ggml-quants.c:390-402); int8 -> half through the biased byte b ^ 0x80 = q + 128
    const __half2 off = __float2half2_rn(1024.0f + 128.0f), zero = __float2half2_rn(0.0f);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        const __half2 s = __float2half2_rn(R.d[2 * cc + hf]);
        const uint4 & a = R.q[4 * cc + 2 * hf], & b = R.q[4 * cc + 2 * hf + 1];
        const uint32_t w[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
#pragma unroll
        for (int i = 0; i < 4; ++i) st_row(a_row + (4 * hf + i) * TC_A_LBO, deq8(w[2 * i] ^ 0x80808080u, w[2 * i + 1] ^ 0x80808080u, off, s, zero));
    }
}

template <int cc> __device__ __forceinline__ void deq_chunk(const RawQ40 & R, int, uint32_t a_row) {
    // block = 16 bytes: low nibbles are weights 0..15, high nibbles 16..31, w = d * (q - 8) (ggml-quants.c:307-325)
    const __half2 off = __float2half2_rn(1024.0f + 8.0f), zero = __float2half2_rn(0.0f);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        const __half2 s = __float2half2_rn(R.d[2 * cc + hf]);
        const uint4 & a = R.q[2 * cc + hf];
        const uint32_t w[4] = { a.x, a.y, a.z, a.w };
#pragma unroll
        for (int i = 0; i < 4; ++i)                                              // i = 0, 1: low nibbles of words 0-1, 2-3; i = 2, 3: high nibbles
            st_row(a_row + (4 * hf + i) * TC_A_LBO, deq8((w[2 * (i & 1)] >> (4 * (i >> 1))) & 0x0f0f0f0fu, (w[2 * (i & 1) + 1] >> (4 * (i >> 1))) & 0x0f0f0f0fu, off, s, zero));
    }
}

// one tile's worth of A stages for thread (row, g): block kb of the row arrives in raw-ring slot (rit + kb) % TC_RAW (TMA, issued by the raw
// producer warp up to TC_RAW blocks ahead); the thread copies its half block to registers, hands the slot back, and fills stage uses
// it0 + 4kb + 2g and + 2g + 1
// d-plane prefetch of the planar types: q6_K one f16 per 256 weights, q8_0 / q4_0 four (this thread's blocks 4g .. 4g+3 of the span's eight)
template <class Raw> struct DAhead { __device__ static void load(const uint8_t *, int64_t, int, float (&)[4]) {} __device__ static void put(Raw &, const float (&)[4]) {} };
template <> struct DAhead<RawQ6K> {
    __device__ static void load(const uint8_t * drow, int64_t kb, int, float (&d)[4]) { d[0] = h2f(__ldg((const uint16_t *) (drow + kb * 2))); }
    __device__ static void put(RawQ6K & R, const float (&d)[4]) { R.d = d[0]; }
};
template <class Raw> struct DAhead4 {
    __device__ static void load(const uint8_t * drow, int64_t kb, int g, float (&d)[4]) {
        const uint2 v = __ldg((const uint2 *) (drow + kb * 16 + g * 8));
        d[0] = h2f(v.x & 0xffff); d[1] = h2f(v.x >> 16); d[2] = h2f(v.y & 0xffff); d[3] = h2f(v.y >> 16);
    }
    __device__ static void put(Raw & R, const float (&d)[4]) { R.d[0] = d[0]; R.d[1] = d[1]; R.d[2] = d[2]; R.d[3] = d[3]; }
};
template <> struct DAhead<RawQ80> : DAhead4<RawQ80> {};
template <> struct DAhead<RawQ40> : DAhead4<RawQ40> {};

template <class Raw>
__device__ __forceinline__ void tc_produce_tile(uint32_t raw_row, uint32_t raw_stride, uint32_t n_raw, const uint8_t * drow, int64_t nkb, int g, int lane, uint32_t it0, uint32_t rit0,
                                                uint32_t stage0, uint32_t a_full, uint32_t empty, uint32_t raw_full, uint32_t raw_empty, int flags) {
    float dn[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
    if (drow) DAhead<Raw>::load(drow, 0, g, dn);                                 // planar types: the f16 d of the span, one span ahead of its use
    for (int64_t kb = 0; kb < nkb; ++kb) {
        const uint32_t ru = rit0 + (uint32_t) kb, r = ru % n_raw;
        Raw R;
        tc_mbar_wait(raw_full + 8 * r, (ru / n_raw) & 1);
        raw_load(R, raw_row + r * raw_stride, g);
        if (drow) { DAhead<Raw>::put(R, dn); if (kb + 1 < nkb) DAhead<Raw>::load(drow, kb + 1, g, dn); }
        // The slot may be overwritten by the next TMA (async proxy) as soon as all 8 warps have arrived, so the generic-proxy ld.shared above must
        // have been PERFORMED, not just issued — their values are not consumed before the arrive, and ptxas schedules SYNCS.ARRIVE right behind
        // the LDS (observed: rows of the next block leaking into this one).  The cross-proxy fence orders them before the TMA write.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(raw_empty + 8 * r);
        const uint32_t it = it0 + (uint32_t) kb * 4 + 2 * g;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const uint32_t u = it + cc, s = u % TC_STAGES;
            tc_mbar_wait(empty + 8 * s, ((u / TC_STAGES) & 1) ^ 1);
            if (!(flags & 1)) { if (cc == 0) deq_chunk<0>(R, g, stage0 + s * (TC_A_BYTES + TC_B_BYTES)); else deq_chunk<1>(R, g, stage0 + s * (TC_A_BYTES + TC_B_BYTES)); }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(a_full + 8 * s);                       // one arrival per warp
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------- activations -> F16 tiles
// X F32 [n, k] (row stride x_ld elements) -> x16 tiles in the canonical UMMA layout; rows n .. n_pad-1 are zero.  8 consecutive threads
// write one 128-byte core matrix.
// k = the K extent of the tile layout (a multiple of 64), k_real <= k = the columns x has (a multiple of 8): chunks past k_real are zero
__global__ void __launch_bounds__(256) k_x_to_f16_tiles(const float * __restrict__ x, int64_t x_ld, uint8_t * __restrict__ x16, int64_t n, int64_t n_pad, int64_t k, int64_t k_real) {
    const int64_t id = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k8n = k >> 3;
    if (id >= n_pad * k8n) return;
    const int r = (int) (id & 7);
    const int64_t k8 = (id >> 3) % k8n, ng = (id >> 3) / k8n;
    const int64_t row = ng * 8 + r;
    uint4 out = make_uint4(0, 0, 0, 0);
    if (row < n && k8 * 8 < k_real) {
        const float4 a = __ldg((const float4 *) (x + row * x_ld + k8 * 8)), b = __ldg((const float4 *) (x + row * x_ld + k8 * 8 + 4));
        __half2 h[4] = { __floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w), __floats2half2_rn(b.x, b.y), __floats2half2_rn(b.z, b.w) };
        out = *(const uint4 *) h;
    }
    const int64_t nt = row / TC_N, nl = row % TC_N, kc64 = k8 >> 3, kc = k8 & 7;
    uint8_t * tile = x16 + (nt * (k >> 6) + kc64) * TC_B_BYTES;
    *(uint4 *) (tile + kc * TC_B_LBO + (nl >> 3) * TC_SBO + (nl & 7) * 16) = out;
}

// ---------------------------------------------------------------------------------------------------------------- the GEMM kernel
__global__ void __launch_bounds__(TC_THREADS, 1) k_mmq_tc(const __grid_constant__ TcArgs A, const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap wmap1,
                                                          const __grid_constant__ CUtensorMap wmap2) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = tc_smem_u32(smem);
    const uint32_t raw0 = sbase + TC_STAGES * (TC_A_BYTES + TC_B_BYTES);       // raw quant-block ring
    const uint32_t bars = raw0 + TC_RAW_REGION;                         // a_full[S], b_full[S], empty[S], tmem_full[2], tmem_empty[2], raw_full[R], raw_empty[R], tmem ptr
    const uint32_t a_full = bars, b_full = bars + 8 * TC_STAGES, empty = bars + 16 * TC_STAGES, t_full = bars + 24 * TC_STAGES, t_empty = t_full + 16;
    const uint32_t raw_full = t_empty + 16, raw_empty = raw_full + 8 * TC_RAW_MAX;
    volatile uint32_t * tmem_slot = (volatile uint32_t *) (smem + TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + TC_RAW_REGION + 24 * TC_STAGES + 32 + 16 * TC_RAW_MAX);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { tc_mbar_init(a_full + 8 * s, TC_DEQ_THREADS / 64); tc_mbar_init(b_full + 8 * s, 1); tc_mbar_init(empty + 8 * s, 1); }
        for (int i = 0; i < 2; ++i) { tc_mbar_init(t_full + 8 * i, 1); tc_mbar_init(t_empty + 8 * i, 128); }
        for (int r = 0; r < TC_RAW_MAX; ++r) { tc_mbar_init(raw_full + 8 * r, 1); tc_mbar_init(raw_empty + 8 * r, TC_DEQ_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32((const void *) tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int n_tiles = A.tiles_m * A.tiles_n * A.splitk, nkc = (int) (A.k >> 6) / A.splitk;        // work items; K chunks of 64 per work item
    uint32_t it = 0;                                                          // stage uses so far (same sequence in every role)

    if (warp < 8) {
        // ================================================================== A producers (dequant)
        const int row = threadIdx.x & 127, g = threadIdx.x >> 7;
        const uint32_t stage0 = sbase + row * 16;                              // row group (row >> 3) * TC_SBO + (row & 7) * 16
        const int64_t nkb = (A.k >> 8) / A.splitk;                            // 256-weight spans per work item
        uint32_t rit = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, it += (uint32_t) nkc, rit += (uint32_t) nkb) {
            const int mtg = (t / A.splitk) % A.tiles_m, si = tc_seg_of(A, mtg), mt = mtg - A.seg[si].tile0;
            const int64_t sm = A.seg[si].m;
            const uint8_t * swd = A.seg[si].wd;
            int64_t gr = (int64_t) mt * TC_M + row; if (gr >= sm) gr = sm - 1;          // tail rows (zero-filled by TMA): any valid d, never stored
            const uint32_t rrow = raw0 + row * A.seg[si].span_bytes, rstride = (uint32_t) A.raw_stride, nr = (uint32_t) A.n_raw;
            switch (A.seg[si].type) {
                case B200_Q4_K: tc_produce_tile<RawQ4K>(rrow, rstride, nr, nullptr, nkb, g, lane, it, rit, stage0, a_full, empty, raw_full, raw_empty, A.flags); break;
                case B200_Q5_K: tc_produce_tile<RawQ5K>(rrow, rstride, nr, nullptr, nkb, g, lane, it, rit, stage0, a_full, empty, raw_full, raw_empty, A.flags); break;
                case B200_Q6_K: tc_produce_tile<RawQ6K>(rrow, rstride, nr, swd + (gr * nkb * A.splitk + (t % A.splitk) * nkb) * 2, nkb, g, lane, it, rit, stage0, a_full, empty, raw_full, raw_empty, A.flags); break;
                case B200_Q8_0: tc_produce_tile<RawQ80>(rrow, rstride, nr, swd + (gr * nkb * A.splitk + (t % A.splitk) * nkb) * 16, nkb, g, lane, it, rit, stage0, a_full, empty, raw_full, raw_empty, A.flags); break;
                default:        tc_produce_tile<RawQ40>(rrow, rstride, nr, swd + (gr * nkb * A.splitk + (t % A.splitk) * nkb) * 16, nkb, g, lane, it, rit, stage0, a_full, empty, raw_full, raw_empty, A.flags); break;
            }
        }
    } else if (warp == 8) {
        // ================================================================== B producer (TMA bulk copies of pre-tiled F16 activations)
        if (lane == 0) {
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int ntx = (t / A.splitk) / A.tiles_m, nt = ntx / A.nsub, half = ntx % A.nsub;
                const uint8_t * src = A.x16 + ((int64_t) nt * nkc * A.splitk + (int64_t) (t % A.splitk) * nkc) * TC_B_BYTES;
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const uint32_t s = it % TC_STAGES;
                    tc_mbar_wait(empty + 8 * s, ((it / TC_STAGES) & 1) ^ 1);
                    if (A.flags & 4) { tc_mbar_arrive(b_full + 8 * s); continue; }
                    const uint32_t dst = sbase + s * (TC_A_BYTES + TC_B_BYTES) + TC_A_BYTES;
                    tc_load_b(dst, src + (int64_t) kc * TC_B_BYTES, A.nsub, half, b_full + 8 * s, true);
                }
            }
        }
    } else if (warp == 9) {
        // ================================================================== MMA issuer
        if (lane == 0) {
            uint32_t tcount = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tcount) {
                const int ntx = (t / A.splitk) / A.tiles_m, wn = TC_N / A.nsub;
                int64_t ncols = A.n - (int64_t) ntx * wn; if (ncols > wn) ncols = wn;
                const uint32_t idesc = tc_idesc((int) ((ncols + 15) & ~15));
                const uint32_t b_lbo = (uint32_t) (wn / 8) * 128;
                const uint64_t b_hi = tc_desc(0, b_lbo, TC_SBO);              // everything but the start address; the K-step advances the address by 2 LBO
                const uint32_t acc = tcount & 1, d_tmem = tmem + acc * TC_N;
                tc_mbar_wait(t_empty + 8 * acc, ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const uint32_t s = it % TC_STAGES, par = (it / TC_STAGES) & 1;
                    tc_mbar_wait(a_full + 8 * s, par);
                    tc_mbar_wait(b_full + 8 * s, par);
                    tc_fence_after();
                    const uint32_t a_s = sbase + s * (TC_A_BYTES + TC_B_BYTES), b_s = a_s + TC_A_BYTES;
#pragma unroll
                    for (int j = 0; j < TC_K / 16; ++j)
                        if (!(A.flags & 2)) tc_mma(d_tmem, tc_desc(a_s + j * 2 * TC_A_LBO, TC_A_LBO, TC_SBO), b_hi | (uint64_t) ((b_s + j * 2 * b_lbo) >> 4), idesc, (kc | j) != 0);
                    tc_commit(empty + 8 * s);                                 // frees the stage when these MMAs have read it
                }
                tc_commit(t_full + 8 * acc);                                  // accumulator complete
            }
        }
    } else if (warp == 14) {
        // ================================================================== raw weight producer: one 2-D TMA (128 rows x one quant block) per 4 stages
        if (lane == 0) {
            const int nkb = (int) (A.k >> 8) / A.splitk;
            const uint32_t n_raw = (uint32_t) A.n_raw, rstride = (uint32_t) A.raw_stride;
            uint32_t ru = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int mtg = (t / A.splitk) % A.tiles_m, si = tc_seg_of(A, mtg), mt = mtg - A.seg[si].tile0;
                const int blk = A.seg[si].span_bytes;
                const CUtensorMap * map = si == 0 ? &wmap : si == 1 ? &wmap1 : &wmap2;
                for (int kb = 0; kb < nkb; ++kb, ++ru) {
                    const uint32_t r = ru % n_raw;
                    tc_mbar_wait(raw_empty + 8 * r, ((ru / n_raw) & 1) ^ 1);
                    tc_mbar_expect_tx(raw_full + 8 * r, (uint32_t) (TC_M * blk));
                    tc_tma_2d(raw0 + r * rstride, map, ((t % A.splitk) * nkb + kb) * blk, mt * TC_M, raw_full + 8 * r);
                }
            }
        }
    } else {
        // ================================================================== epilogue (warps 10..13 -> TMEM lane quarters 2, 3, 0, 1)
        const int quarter = warp & 3;
        uint32_t tcount = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tcount) {
            const int mtg = (t / A.splitk) % A.tiles_m, ntx = (t / A.splitk) / A.tiles_m, wn = TC_N / A.nsub;
            const int si = tc_seg_of(A, mtg), mt = mtg - A.seg[si].tile0;
            const int64_t seg_m = A.seg[si].m, seg_ld = A.seg[si].dst_ld;
            const uint32_t acc = tcount & 1;
            tc_mbar_wait(t_full + 8 * acc, (tcount >> 1) & 1);
            tc_fence_after();
            const int64_t mrow = (int64_t) mt * TC_M + quarter * 32 + lane;
            int64_t ncols = A.n - (int64_t) ntx * wn; if (ncols > wn) ncols = wn;
            float * out = A.seg[si].dst + ((int64_t) ntx * wn) * seg_ld + mrow;
            const bool fuse_add = mrow < seg_m && A.splitk == 1 && A.resid;  // fused residual ADD (wo / ffn_down of a llama-family layer): one F32 add, as the separate op would do (single-matrix launches only)
            const float * rs = A.resid + ((int64_t) ntx * wn) * seg_ld + mrow;
            for (int cc = 0; cc * 32 < ncols; ++cc) {
                uint32_t v[32];
                float r[32];
                // the residual may BE the destination, so the compiler cannot move these loads above the stores below by itself: all 32 are issued up front
                // (they do not depend on the accumulator and overlap the TMEM read)
                if (fuse_add) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = cc * 32 + j < ncols ? __ldcg(rs + (int64_t) (cc * 32 + j) * seg_ld) : 0.0f;
                }
                tc_ld32(tmem + ((uint32_t) (quarter * 32) << 16) + acc * TC_N + cc * 32, v);
                if (fuse_add) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (cc * 32 + j < ncols) out[(int64_t) (cc * 32 + j) * seg_ld] = __fadd_rn(__uint_as_float(v[j]), r[j]);
                } else if (mrow < seg_m && A.splitk == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (cc * 32 + j < ncols) out[(int64_t) (cc * 32 + j) * seg_ld] = __uint_as_float(v[j]);
                } else if (mrow < seg_m) {
                    for (int j = 0; j < 32; ++j) if (cc * 32 + j < ncols) atomicAdd(out + (int64_t) (cc * 32 + j) * seg_ld, __uint_as_float(v[j]));
                }
            }
            tc_fence_before();
            tc_mbar_arrive(t_empty + 8 * acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}


// ================================================================================================================ F16 weights
// Y[m, n] = W[m, k] . X[n, k]^T for F16 W (the F16 model of BASELINE.json configs[4]; the VPM / APM / TTS graphs of SURVEY.md 8f): no dequantisation,
// so the weight tile goes from global memory straight into the tensor core's operand layout — one 2-D TMA per stage with the 128-byte swizzle
// (box = 64 halves x 128 rows = the canonical K-major SWIZZLE_128B UMMA tile, 8-row groups 1024 B apart), activations as above.  6 warps:
// TMA producer, MMA issuer, 4 epilogue warps.  Replaces the cuBLAS branch of ggml_cuda_mul_mat (ggml-cuda.cu:1983, 2001-2084) and mmvf for n > 8;
// arithmetic = the CPU oracle's (activations rounded to F16 = vec_dot_type of F16 weights, ggml-cpu.c:196-350; F32 accumulation).
constexpr int TF_STAGES = 4, TF_THREADS = 192;
constexpr int TF_SMEM = TF_STAGES * (TC_A_BYTES + TC_B_BYTES) + 1024 + 256;

struct TcF16Args { const uint8_t * x16; float * dst; int64_t dst_ld, m, k, n; int tiles_m, tiles_n, nsub, splitk; };

__device__ __forceinline__ uint64_t tc_desc_sw128(uint32_t saddr) {      // K-major SWIZZLE_128B: SBO = 1024 B (8 rows x 128 B), LBO unused, layout type 2
    return (uint64_t) ((saddr & 0x3ffff) >> 4) | ((uint64_t) 1 << 16) | ((uint64_t) (1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(TF_THREADS, 1) k_mm_f16_tc(const __grid_constant__ TcF16Args A, const __grid_constant__ CUtensorMap wmap) {
    extern __shared__ __align__(1024) uint8_t smem_f[];
    const uint32_t sbase = (tc_smem_u32(smem_f) + 1023u) & ~1023u;              // SWIZZLE_128B tiles must sit on 1024-byte boundaries
    const uint32_t bars = sbase + TF_STAGES * (TC_A_BYTES + TC_B_BYTES);
    const uint32_t full = bars, empty = bars + 8 * TF_STAGES, t_full = bars + 16 * TF_STAGES, t_empty = t_full + 16, tmem_slot_a = t_empty + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TF_STAGES; ++s) { tc_mbar_init(full + 8 * s, 1); tc_mbar_init(empty + 8 * s, 1); }
        for (int i = 0; i < 2; ++i) { tc_mbar_init(t_full + 8 * i, 1); tc_mbar_init(t_empty + 8 * i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot_a), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot_a));
    const int n_tiles = A.tiles_m * A.tiles_n * A.splitk, nkc = (int) (A.k >> 6) / A.splitk;
    uint32_t it = 0;
    if (warp == 0) {
        if (lane == 0) {
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int mt = (t / A.splitk) % A.tiles_m, ntx = (t / A.splitk) / A.tiles_m, nt = ntx / A.nsub, half = ntx % A.nsub;
                const uint8_t * src = A.x16 + ((int64_t) nt * nkc * A.splitk + (int64_t) (t % A.splitk) * nkc) * TC_B_BYTES;
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const uint32_t s = it % TF_STAGES;
                    tc_mbar_wait(empty + 8 * s, ((it / TF_STAGES) & 1) ^ 1);
                    tc_mbar_expect_tx(full + 8 * s, TC_A_BYTES + TC_B_BYTES / A.nsub);
                    const uint32_t a_s = sbase + s * (TC_A_BYTES + TC_B_BYTES);
                    tc_tma_2d(a_s, &wmap, ((t % A.splitk) * nkc + kc) * TC_K, mt * TC_M, full + 8 * s);
                    tc_load_b(a_s + TC_A_BYTES, src + (int64_t) kc * TC_B_BYTES, A.nsub, half, full + 8 * s, false);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t tcount = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tcount) {
                const int ntx = (t / A.splitk) / A.tiles_m, wn = TC_N / A.nsub;
                int64_t ncols = A.n - (int64_t) ntx * wn; if (ncols > wn) ncols = wn;
                const uint32_t idesc = tc_idesc((int) ((ncols + 15) & ~15));
                const uint32_t b_lbo = (uint32_t) (wn / 8) * 128;
                const uint64_t b_hi = tc_desc(0, b_lbo, TC_SBO);              // everything but the start address; the K-step advances the address by 2 LBO
                const uint32_t acc = tcount & 1, d_tmem = tmem + acc * TC_N;
                tc_mbar_wait(t_empty + 8 * acc, ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const uint32_t s = it % TF_STAGES;
                    tc_mbar_wait(full + 8 * s, (it / TF_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_s = sbase + s * (TC_A_BYTES + TC_B_BYTES), b_s = a_s + TC_A_BYTES;
#pragma unroll
                    for (int j = 0; j < TC_K / 16; ++j)
                        tc_mma(d_tmem, tc_desc_sw128(a_s + j * 32), b_hi | (uint64_t) ((b_s + j * 2 * b_lbo) >> 4), idesc, (kc | j) != 0);
                    tc_commit(empty + 8 * s);
                }
                tc_commit(t_full + 8 * acc);
            }
        }
    } else {
        const int quarter = warp & 3;
        uint32_t tcount = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tcount) {
            const int mt = (t / A.splitk) % A.tiles_m, ntx = (t / A.splitk) / A.tiles_m, wn = TC_N / A.nsub;
            const uint32_t acc = tcount & 1;
            tc_mbar_wait(t_full + 8 * acc, (tcount >> 1) & 1);
            tc_fence_after();
            const int64_t mrow = (int64_t) mt * TC_M + quarter * 32 + lane;
            int64_t ncols = A.n - (int64_t) ntx * wn; if (ncols > wn) ncols = wn;
            float * out = A.dst + ((int64_t) ntx * wn) * A.dst_ld + mrow;
            for (int cc = 0; cc * 32 < ncols; ++cc) {
                uint32_t v[32];
                tc_ld32(tmem + ((uint32_t) (quarter * 32) << 16) + acc * TC_N + cc * 32, v);
                if (mrow < A.m && A.splitk == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (cc * 32 + j < ncols) out[(int64_t) (cc * 32 + j) * A.dst_ld] = __uint_as_float(v[j]);
                } else if (mrow < A.m) {
                    for (int j = 0; j < 32; ++j) if (cc * 32 + j < ncols) atomicAdd(out + (int64_t) (cc * 32 + j) * A.dst_ld, __uint_as_float(v[j]));
                }
            }
            tc_fence_before();
            tc_mbar_arrive(t_empty + 8 * acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------- host side
bool mmq_tc_supported(int type, int layout, int64_t k, int64_t n, const void * w, int64_t row_stride) {
    if (n <= 8 || k % 256) return false;
    if ((uintptr_t) w % 16) return false;
    if (type == B200_Q4_K || type == B200_Q5_K) return row_stride == k / 256 * type_size(type);             // native 16-byte-multiple blocks
    if (type == B200_Q6_K || type == B200_Q8_0 || type == B200_Q4_0) return layout == B200_LAYOUT_PLANAR;   // payload plane + f16 d plane
    return false;
}
// 128-column tiles when the work items (tiles x split-K) cannot give even half of the SMs something to do (small m, small ubatches)
static int tc_nsub(int tiles_m, int64_t n) {
    static const int thr = getenv("B200_TC_N128_MAX") ? atoi(getenv("B200_TC_N128_MAX")) : 0;
    const int64_t tiles256 = (int64_t) tiles_m * ((n + TC_N - 1) / TC_N);
    return n > TC_N / 2 && tiles256 * 2 <= (thr > 0 ? thr : sm_count()) ? 2 : 1;        // only while the doubled tile count still fits one wave
}
// two CTAs per tile, half of K each, when even the 256-column tiling leaves more than half of the SMs without work (`units` = K in units that
// must stay whole per CTA: 256-weight spans for quantised weights, 128 for F16)
static int tc_splitk(int tiles_m, int64_t n, int64_t units) {
    static const int off = getenv("B200_TC_NO_SPLITK") ? atoi(getenv("B200_TC_NO_SPLITK")) : 0;
    const int64_t tiles256 = (int64_t) tiles_m * ((n + TC_N - 1) / TC_N);
    return !off && units % 2 == 0 && units >= 4 && tiles256 * 2 <= sm_count() ? 2 : 1;
}
// will mmq_tc carry a residual in its epilogue for this shape?  (not under split-K: three addends would not commute)
bool mmq_tc_fuses_resid(int64_t m, int64_t k, int64_t n) { return tc_splitk((int) ((m + TC_M - 1) / TC_M), n, k / 256) == 1; }
size_t mmq_tc_scratch_bytes(int64_t k, int64_t n) { return (size_t) ((n + TC_N - 1) / TC_N * TC_N) * (size_t) ((k + TC_K - 1) / TC_K * TC_K) * 2; }

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda)
typedef CUresult (*tc_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tc_encode_fn tc_encoder() {
    static tc_encode_fn fn = [] {
        void * p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (tc_encode_fn) p;
    }();
    return fn;
}

static int tc_make_map(CUtensorMap & wmap, const void * w, int type, int64_t m, int64_t k) {
    // the payload plane as a 2-D byte tensor [m rows][row bytes]; box = one quant block x 128 rows
    const cuuint64_t blk = (cuuint64_t) (256 / blck_size(type)) * payload_size(type), rowb = (cuuint64_t) (k / 256) * blk;
    const cuuint64_t gdim[2] = { rowb, (cuuint64_t) m }, gstr[1] = { rowb };
    const cuuint32_t box[2] = { (cuuint32_t) blk, TC_M }, estr[2] = { 1, 1 };
    static const bool nopromo = getenv("B200_TC_NOPROMO") != nullptr;
    return tc_encoder()(&wmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *) w, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        (nopromo ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
           ? B200_OK : B200_ERR_UNSUPPORTED;
}

// can these matrices (all [m_i, k], the same activations) run as ONE launch?  K-quants only (one raw-ring geometry), and only when the merged tile count needs no split-K
bool mmq_tc_multi_ok(int nseg, const int * type, const int64_t * m, int64_t k, int64_t n) {
    if (nseg < 2 || nseg > 3) return false;
    int64_t tiles_m = 0;
    for (int i = 0; i < nseg; ++i) { if (!is_kquant(type[i]) || m[i] <= 0) return false; tiles_m += (m[i] + TC_M - 1) / TC_M; }
    return tc_splitk((int) tiles_m, n, k / 256) == 1;
}

int mmq_tc_multi(int nseg, const void * const * w, const int * type, const int64_t * m, int64_t k, const float * x, int64_t x_ld, int64_t n, float * const * dst, const int64_t * dst_ld,
                 void * scratch, bool reuse_tiles, cudaStream_t st, const float * resid, bool * resid_fused) {
    if (resid_fused) *resid_fused = false;
    static smem_mask_t done{0};
    B200_CUDA_TRY(ensure_dyn_smem(k_mmq_tc, TC_SMEM, done));
    if (!tc_encoder() || nseg < 1 || nseg > 3) return B200_ERR_UNSUPPORTED;
    CUtensorMap wmap[3];
    for (int i = 0; i < 3; ++i) {                                           // one encode per weight matrix (~2 us of host time each); unused slots repeat segment 0
        if (i < nseg) { if (tc_make_map(wmap[i], w[i], type[i], m[i], k)) return B200_ERR_UNSUPPORTED; }
        else wmap[i] = wmap[0];
    }
    const int64_t n_pad = (n + TC_N - 1) / TC_N * TC_N, threads = n_pad * (k >> 3);
    if (!reuse_tiles) {
        k_x_to_f16_tiles<<<(unsigned) ((threads + 255) / 256), 256, 0, st>>>(x, x_ld, (uint8_t *) scratch, n, n_pad, k, k);
        B200_LAUNCH_CHECK();
    }
    TcArgs A = {};
    static const int env_flags = getenv("B200_TC_FLAGS") ? atoi(getenv("B200_TC_FLAGS")) : 0;
    A.flags = env_flags;
    const int64_t nkb = k / 256;
    A.w = (const uint8_t *) w[0]; A.type = type[0]; A.m = m[0]; A.k = k; A.n = n; A.dst = dst[0]; A.dst_ld = dst_ld[0]; A.x16 = (const uint8_t *) scratch;
    A.nseg = nseg; A.tiles_m = 0; A.n_raw = TC_RAW_MAX; A.raw_stride = 0;
    for (int i = 0; i < nseg; ++i) {
        TcArgs::Seg & S = A.seg[i];
        S.type = type[i]; S.m = m[i]; S.dst = dst[i]; S.dst_ld = dst_ld[i]; S.tile0 = A.tiles_m;
        S.span_bytes = (256 / blck_size(type[i])) * payload_size(type[i]);       // 144, 176, 208, 256 (q8_0), 128 (q4_0)
        S.wd = payload_size(type[i]) != type_size(type[i]) ? (const uint8_t *) w[i] + m[i] * nkb * S.span_bytes : nullptr;     // planar: f16 d plane behind the payload plane
        const int nr = TC_RAW_REGION / (TC_M * S.span_bytes) >= TC_RAW_MAX ? TC_RAW_MAX : TC_RAW_REGION / (TC_M * S.span_bytes);
        if (nr < A.n_raw) A.n_raw = nr;
        if (TC_M * S.span_bytes > A.raw_stride) A.raw_stride = TC_M * S.span_bytes;
        A.tiles_m += (int) ((m[i] + TC_M - 1) / TC_M);
    }
    A.span_bytes = A.seg[0].span_bytes; A.row_bytes = nkb * A.span_bytes; A.wd = A.seg[0].wd;
    A.splitk = tc_splitk(A.tiles_m, n, k / 256);
    if (nseg > 1 && A.splitk != 1) return B200_ERR_UNSUPPORTED;                 // (mmq_tc_multi_ok says so beforehand)
    A.resid = A.splitk == 1 && nseg == 1 ? resid : nullptr;                     // three addends would not commute: split-K leaves the ADD to the caller
    if (resid_fused) *resid_fused = A.resid != nullptr;
    A.nsub = tc_nsub(A.tiles_m * A.splitk, n); A.tiles_n = (int) ((n + TC_N / A.nsub - 1) / (TC_N / A.nsub));
    if (A.splitk > 1) B200_CUDA_TRY(cudaMemset2DAsync(dst[0], (size_t) dst_ld[0] * 4, 0, (size_t) m[0] * 4, (size_t) n, st));
    int grid = A.tiles_m * A.tiles_n * A.splitk; if (grid > sm_count()) grid = sm_count();
    k_mmq_tc<<<grid, TC_THREADS, TC_SMEM, st>>>(A, wmap[0], wmap[1], wmap[2]);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int mmq_tc(const void * w, int type, int64_t m, int64_t k, const float * x, int64_t x_ld, int64_t n, float * dst, int64_t dst_ld, void * scratch, bool reuse_tiles, cudaStream_t st,
           const float * resid, bool * resid_fused) {
    return mmq_tc_multi(1, &w, &type, &m, k, x, x_ld, n, &dst, &dst_ld, scratch, reuse_tiles, st, resid, resid_fused);
}

bool mm_f16_tc_supported(int type, int64_t k, int64_t n, const void * w, int64_t row_stride) {
    // k need not be a multiple of the 64-wide K tile (SigLip's ffn_down has k = 4304): the weight tile's tail is zero-filled by TMA (the tensor map knows the real k), the
    // activation tiles' tail by k_x_to_f16_tiles; rows must still be 16-byte multiples for the tensor map's global stride
    return type == B200_F16 && n > 8 && k >= 8 && k % 8 == 0 && (uintptr_t) w % 16 == 0 && row_stride == k * 2;
}

int mm_f16_tc(const void * w, int64_t m, int64_t k, const float * x, int64_t x_ld, int64_t n, float * dst, int64_t dst_ld, void * scratch, bool reuse_tiles, cudaStream_t st) {
    static smem_mask_t done{0};
    B200_CUDA_TRY(ensure_dyn_smem(k_mm_f16_tc, TF_SMEM, done));
    if (!tc_encoder()) return B200_ERR_UNSUPPORTED;
    CUtensorMap wmap;
    {
        const cuuint64_t gdim[2] = { (cuuint64_t) k, (cuuint64_t) m }, gstr[1] = { (cuuint64_t) k * 2 };
        const cuuint32_t box[2] = { TC_K, TC_M }, estr[2] = { 1, 1 };
        if (tc_encoder()(&wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *) w, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return B200_ERR_UNSUPPORTED;
    }
    const int64_t k_pad = (k + TC_K - 1) / TC_K * TC_K, kt = k_pad / TC_K;
    const int64_t n_pad = (n + TC_N - 1) / TC_N * TC_N, threads = n_pad * (k_pad >> 3);
    if (!reuse_tiles) {
        k_x_to_f16_tiles<<<(unsigned) ((threads + 255) / 256), 256, 0, st>>>(x, x_ld, (uint8_t *) scratch, n, n_pad, k_pad, k);
        B200_LAUNCH_CHECK();
    }
    TcF16Args A = {};
    A.x16 = (const uint8_t *) scratch; A.dst = dst; A.dst_ld = dst_ld; A.m = m; A.k = k_pad; A.n = n;
    A.tiles_m = (int) ((m + TC_M - 1) / TC_M);
    // split-K halves the K TILES between two CTAs: only an even tile count splits (k / 128 used to round an odd count down — k = 576 has 9 tiles — and the second
    // CTA would have stopped one tile short)
    A.splitk = tc_splitk(A.tiles_m, n, kt % 2 == 0 ? kt / 2 : 1);
    A.nsub = tc_nsub(A.tiles_m * A.splitk, n); A.tiles_n = (int) ((n + TC_N / A.nsub - 1) / (TC_N / A.nsub));
    if (A.splitk > 1) B200_CUDA_TRY(cudaMemset2DAsync(dst, (size_t) dst_ld * 4, 0, (size_t) m * 4, (size_t) n, st));
    int grid = A.tiles_m * A.tiles_n * A.splitk; if (grid > sm_count()) grid = sm_count();
    k_mm_f16_tc<<<grid, TF_THREADS, TF_SMEM, st>>>(A, wmap);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

} // namespace b200
