// stream_decode.cuh — phase descriptors of the persistent decode kernel (stream_decode.cu).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int SD_WARPS       = 12;                                 // CONSUMER warps (384 threads); one more warp is the producer -> 416 threads, 152 registers each
constexpr int SD_THREADS     = SD_WARPS * 32;
constexpr int SD_DEPTH       = 3;                                  // ring slots per warp: 1 being computed + 2 in flight
constexpr int SD_SLOT_BYTES  = 4608;                               // one unit: 2 q4_K rows of k=4096, 1 q6_K row, or 1 gate/up pair
constexpr int SD_RING_BYTES  = SD_WARPS * SD_DEPTH * SD_SLOT_BYTES;   // 162 KB of weights in flight per SM
constexpr int SD_ACT_BYTES   = 16 * 1024;                          // q8 activation record, k <= 12288
constexpr int SD_ATTN_BYTES  = 32 * 1024;                          // attention phase: cross-warp reduction of the split-KV partials
constexpr int SD_STASH_ROWS  = 128;                                // rows of one phase a CTA may own when it stashes a residual / K-splits inside the CTA
constexpr int SD_KSL_MAX     = 6;                                  // K-slices inside a CTA (must divide SD_WARPS)
constexpr int SD_NPH         = 6;                                  // staged phase descriptors (ring in shared memory)
constexpr int SD_INFLIGHT    = 3;                                  // bulk copies on the wire per consumer ring (pacing, see sd_producer)
constexpr int SD_STAGE_AHEAD = 2;                                  // the producer stages this many phases ahead of the one it issues

enum { SD_MATVEC = 0, SD_ATTN = 1 };
enum { SD_EPI_STORE = 0, SD_EPI_SWIGLU = 1 };
enum { SD_PRO_ACT = 0, SD_PRO_RMSNORM_QUANT = 1, SD_PRO_QUANT = 2 };

struct SdMat {
    const uint8_t * payload; const uint8_t * dplane;      // dplane: planar f16 d plane (q4_0 / q8_0 / q6_K); null for q4_K / q5_K
    float * y; const float * residual;
    int32_t type, rows, row_bytes_p, row_bytes_d;         // bytes of one FULL row in each plane (= row stride)
};

struct SdAttn {                                           // batch-1 attention over the F16 KV cache of one layer (Qwen3 layout)
    const float * q; const float * k_new; const float * v_new;   // raw wq / wk / wv outputs of this token (TAGGED pairs, see SdPhase)
    const float * q_norm_w; const float * k_norm_w;       // [D] or null (no q/k norm: llama arch)
    uint8_t * k_cache; uint8_t * v_cache;                 // F16 [n_ctx][n_head_kv * D]
    int64_t k_row_bytes, v_row_bytes;
    float * out;                                          // [n_head * D] attention output (TAGGED pairs)
    int32_t in_tag, pad_;                                 // phase index + 1 of the matvec phase that produced q / k_new / v_new
    float * part_acc; float2 * part_ml; unsigned * tickets;   // split-KV partials + per-kv-head arrival counters (zero between uses)
    int32_t n_head, n_head_kv, head_dim, rope_mode;
    float scale, eps, theta_scale, freq_scale, ext_factor, attn_factor, corr0, corr1;
};

// Inter-phase hand-off WITHOUT grid barriers ("flag in data"): every vector one phase hands to a later one is an array of 8-byte pairs
// { f32 value, u32 tag }, tag = launch epoch + (producing phase index + 1), written with ONE 8-byte store per element.  A consumer simply loads
// the pairs and re-loads until every tag matches: the poll IS the data load, so a transition costs one store -> L2 -> load trip instead of
// store, fence, atomic arrive, acquire poll, exit broadcast, load.  No fences are needed (8-byte single-copy atomicity carries value and tag
// together) and nothing else crosses CTAs un-tagged (the attention split-KV partials keep their fence + ticket).  Buffers are re-used every
// layer (5 phases later); a CTA that writes phase p+5 has consumed every CTA's phase p+4 output, so all CTAs are past the readers of phase p.
struct alignas(16) SdPhase {
    int32_t kind, n_mat, epilogue, prologue, k, act_group;
    int32_t next_kind, next_mv;                           // kind of phase p+1 (-1: none) and index of the next matvec phase (-1: none)
    int32_t y_tagged;                                     // matvec outputs are stored as tagged pairs (mat[].y then counts PAIRS)
    int32_t ksl;                                          // > 1: K-split INSIDE the CTA — warp w reduces over K-slice w % ksl (12 % ksl == 0); the slices of a
                                                          // row are summed, in slice order, through shared memory at the end of the phase (+ the stashed residual)
    float eps; int32_t ksplit;                            // ksplit > 1: CTA c reduces over K-slice c % ksplit and stores a PARTIAL y
    const float * x[4]; int32_t n_x;                      // prologue input = x[0] + x[1] + ... (fixed order: deterministic)
    int32_t x_tag;                                        // 0: x[] are plain f32 vectors; else x[0] (n_x == 1) is a tagged vector produced by phase x_tag - 1
    int32_t stash_T, resid_stash;                         // stash_T > 0: the prologue keeps its summed input rows [T*c/n, T*(c+1)/n) of THIS CTA in shared memory; a
                                                          // later phase with resid_stash = 1 (T = its row count) adds them as the residual in its epilogue
    float * x_out;                                        // optional: CTA 0 stores the summed input (the new residual stream)
    const float * norm_w; const uint8_t * act; float * norm_out;
    int64_t y_part_stride;                                // ksplit > 1: partial of slice s goes to y + s * y_part_stride
    SdMat mat[3];
    SdAttn attn;
};

struct SdRuntime {                                        // per-step inputs (device pointers; contents change every token)
    const int32_t * pos; const int64_t * kv_idx; const __half * mask; int32_t n_kv;
    float theta_scale, freq_scale, ext_factor, attn_factor, corr0, corr1; int32_t rope_mode, has_rope;   // RoPE of the token (all layers)
    int32_t flags, pad_;
    unsigned long long * prof;                            // optional [n_phases][4] globaltimer stamps of CTA 0: start, prologue done, work done, barrier done
};

struct SegTab {                                           // a CTA's share of one matvec phase (see stream_decode.cu "geometry")
    int nrows[3], rpu[3], upre[4];                        // per matrix: rows of this CTA, rows per unit; unit prefix sums
    int kpart, nm, ksplit, ksl;                           // K-slice of this CTA; matrices per unit (2 = gate/up pair); K-slices inside the CTA
    int nunits, pad_[3];                                  // units of this CTA = row units x ksl
    int sub_p[3], sub_d[3], rbp[3], rbd[3], type[3];      // bytes of one row's K-slice / of one full row, per plane; weight type
    const uint8_t * pay[3], * dpl[3], * pay2, * dpl2;     // first row's slice of this CTA in each plane (pay2/dpl2: the `up` matrix)
    float * y[3]; const float * resid[3];                 // output / residual at this CTA's first row
};

constexpr int SD_SMEM_BYTES = 227 * 1024;                         // the whole opt-in maximum; the map is laid out (and checked) in stream_decode.cu
static_assert(SD_SMEM_BYTES <= 227 * 1024, "k_stream shared memory");

} // namespace b200
