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
constexpr int SD_ATTN_BYTES  = 40 * 1024;                          // attention phase: cross-warp reduction of the split-KV partials
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
    const float * q; const float * k_new; const float * v_new;   // raw wq / wk / wv outputs of this token
    const float * q_norm_w; const float * k_norm_w;       // [D] or null (no q/k norm: llama arch)
    uint8_t * k_cache; uint8_t * v_cache;                 // F16 [n_ctx][n_head_kv * D]
    int64_t k_row_bytes, v_row_bytes;
    float * out;                                          // [n_head * D] attention output (F32)
    float * part_acc; float2 * part_ml; unsigned * tickets;   // split-KV partials + per-kv-head arrival counters (zero between uses)
    int32_t n_head, n_head_kv, head_dim, rope_mode;
    float scale, eps, theta_scale, freq_scale, ext_factor, attn_factor, corr0, corr1;
};

struct alignas(16) SdPhase {
    int32_t kind, n_mat, epilogue, prologue, k, act_group;
    int32_t next_kind, next_mv, pad1_, pad2_;             // kind of phase p+1 (-1: none) and index of the next matvec phase (-1: none)
    float eps; int32_t ksplit;                            // ksplit > 1: CTA c reduces over K-slice c % ksplit and stores a PARTIAL y
    const float * x[4]; int32_t n_x, pad0_;               // prologue input = x[0] + x[1] + ... (fixed order: deterministic)
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
    int kpart, nm, ksplit, pad_;                          // K-slice of this CTA; matrices per unit (2 = gate/up pair)
    int sub_p[3], sub_d[3], rbp[3], rbd[3], type[3];      // bytes of one row's K-slice / of one full row, per plane; weight type
    const uint8_t * pay[3], * dpl[3], * pay2, * dpl2;     // first row's slice of this CTA in each plane (pay2/dpl2: the `up` matrix)
    float * y[3]; const float * resid[3];                 // output / residual at this CTA's first row
};

constexpr int SD_SMEM_BYTES = SD_RING_BYTES + SD_ACT_BYTES + SD_ATTN_BYTES + 2 * SD_WARPS * SD_DEPTH * 8 /* full + empty barriers */
                            + SD_WARPS * SD_DEPTH * 32 /* slot descriptors */ + 64 * 4 + SD_NPH * (int) (sizeof(SdPhase) + sizeof(SegTab)) + 64 + 64 * 8 + 16;
static_assert(SD_SMEM_BYTES <= 227 * 1024, "k_stream shared memory");

} // namespace b200
