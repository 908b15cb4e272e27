/* include/b200_ops.h — C-ABI of the hand-written sm_100a kernels (libb200ops.so).
 *
 * This is the "thin C-ABI of hand-written sm_100a kernels" the host side calls (BASELINE.json north_star): plain
 * pointers, sizes and strides, no C++/torch/ggml types.  The ggml backend plugin (include/ggml-b200.h,
 * libggml-b200.so) maps each ggml op onto one of these entry points; parity tests call them through ctypes.
 *
 * Every entry point cites the reference CUDA-backend interface it replaces (paths relative to the reference root,
 * ggml/src/ggml-cuda/...) and follows the arithmetic of the reference CPU backend (the parity oracle, oracle/*.c).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless named host_*; `stream` is a cudaStream_t passed as void*.
 *   - return value: 0 = launched OK; B200_ERR_UNSUPPORTED = shape/type outside what the kernel handles (caller must not
 *     have routed it here: the plugin's supports_op mirrors b200_*_supported); other negative = CUDA error code negated.
 *   - tensors follow ggml's convention: ne[0] is the contiguous dimension, nb[] are BYTE strides.
 */
#ifndef B200_OPS_H
#define B200_OPS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_API __attribute__((visibility("default")))

#define B200_OK               0
#define B200_ERR_UNSUPPORTED (-1000)
#define B200_ERR_ARG         (-1001)

/* ggml_type ids this library understands (ggml/include/ggml.h:379-421) */
enum b200_type {
    B200_F32 = 0, B200_F16 = 1, B200_Q4_0 = 2, B200_Q8_0 = 8, B200_Q4_K = 12, B200_Q5_K = 13, B200_Q6_K = 14,
    B200_I32 = 26, B200_I64 = 27, B200_BF16 = 30,
};

/* Strided 4-D tensor view: the fields of struct ggml_tensor (ggml.h:626-658) a kernel needs. */
typedef struct b200_tensor {
    void *  data;
    int32_t type;      /* enum b200_type */
    int32_t layout;    /* B200_LAYOUT_*: only quantised weights can be non-native */
    int64_t ne[4];
    int64_t nb[4];
} b200_tensor;

/* Weight layouts in HBM.  NATIVE = ggml block structs back to back (ggml-common.h).  PLANAR = the same bytes split
 * into a 16-byte aligned payload plane followed by a scale plane, so that q4_0 / q8_0 / q6_K rows can be streamed with
 * 128-bit loads (their native blocks are 18 / 34 / 210 bytes).  See DESIGN.md "Data layout in HBM".
 *   q4_0 : payload 16 B (qs)                  | scale plane: f16 d
 *   q8_0 : payload 32 B (qs)                  | scale plane: f16 d
 *   q6_K : payload 208 B (ql, qh, scales)     | scale plane: f16 d
 * q4_K (144 B) and q5_K (176 B) are already 16-byte multiples and are never repacked. */
enum { B200_LAYOUT_NATIVE = 0, B200_LAYOUT_PLANAR = 1 };

/* ---- library / device --------------------------------------------------------------------------------------------- */
B200_API int          b200_abi_version(void);                      /* bumps when any signature below changes */
B200_API const char * b200_error_string(int code);
B200_API int          b200_device_sm_count(int device);

/* ---- weight repack (buffer set_tensor / get_tensor path; replaces nothing in the reference: ggml-cuda keeps native
 *      blocks and pays for it with 2-/4-byte loads, vecdotq.cuh:103-122,580-600) ---------------------------------------
 * Scatter `size` bytes that sit at native byte offset `offset` of a [nblocks] block array into the planar layout at
 * `dst_planar` (and the inverse).  Any offset/size works (llama-model-loader.cpp:1077-1093 uploads 1 MiB chunks). */
B200_API int b200_repack_supported(int type);
B200_API int b200_repack_scatter(int type, const void * src_native_chunk, void * dst_planar, int64_t nblocks_total,
                                 int64_t offset, int64_t size, void * stream);
B200_API int b200_repack_gather (int type, const void * src_planar, void * dst_native_chunk, int64_t nblocks_total,
                                 int64_t offset, int64_t size, void * stream);

/* ---- activation quantisation (replaces quantize_q8_1, ggml-cuda/quantize.cu:4-48; arithmetic of the CPU oracle:
 *      quantize_row_q8_K_ref ggml-quants.c:2555-2592 for K-quant weights, quantize_row_q8_0 for q4_0/q8_0) ---------------
 * Device-side activation buffer ("act"), one record per column, planar:
 *     int8  qs[k]            quantised values
 *     float d[k/G]           G = 256 for K-quant weights (q8_K), 32 for q4_0/q8_0 weights (q8_0; value is f16-rounded)
 *     int16 bsum[k/16|k/32]  partial sums of qs (16-wide for G = 256, 32-wide for G = 32)
 * b200_act_bytes gives the size of one record; records are laid out back to back, 16-byte aligned. */
B200_API size_t b200_act_bytes(int weight_type, int64_t k);
B200_API int    b200_quantize_act(int weight_type, const float * x, int64_t x_col_stride_elems, void * act,
                                  int64_t k, int64_t ncols, void * stream);

/* ---- MUL_MAT (replaces ggml_cuda_mul_mat ggml-cuda.cu:2001-2084: mmvq.cu mul_mat_vec_q for n <= 8, mmq.cu mul_mat_q
 *      beyond; oracle: ggml_compute_forward_mul_mat ggml-cpu.c:1210-1402) ------------------------------------------------
 * dst[m, n] (F32) = W[m, k] (any supported type) . X[n, k]^T (F32; F16 when W is F16 — the im2col x kernel product of ggml_conv_1d / conv_2d, the one
 * F16-activation MUL_MAT the CPU backend has), batched over ne[2], ne[3] with ggml broadcast rules.
 * `scratch` must hold b200_mul_mat_scratch_bytes(...) bytes (activation records / tile buffers).
 * Routing: n <= 8 columns -> dequant-in-register matvec on q8_K / q8_0 activation records (the CPU oracle's integer arithmetic);
 *          n >  8 columns, q4_K / q5_K native or q6_K / q8_0 / q4_0 planar, k % 256 == 0 -> tcgen05 dequant-GEMM k_mmq_tc (csrc/mmq_tc.cu:
 *          F16 operands, F32 accumulation in TMEM); F16 weights, n > 8, k % 8 == 0 -> k_mm_f16_tc (TMA-fed tcgen05 GEMM; a K tail is zero-filled);
 *          other float-weight products with n > 8 (F32 x F32, BF16, unaligned rows), many small batch slices, F16 activations -> k_mm_simt (tiled F32 GEMM,
 *          one launch over all batch slices; replaces the cuBLAS routes of ggml_cuda_mul_mat); everything else -> column-chunked matvec / warp-per-row float kernel. */
B200_API int    b200_mul_mat_supported(const b200_tensor * w, const b200_tensor * x, const b200_tensor * dst);
B200_API size_t b200_mul_mat_scratch_bytes(const b200_tensor * w, const b200_tensor * x);
B200_API int    b200_mul_mat(const b200_tensor * w, const b200_tensor * x, const b200_tensor * dst, void * scratch,
                             size_t scratch_bytes, void * stream);
/* Same, with flags.  B200_MM_REUSE_ACT: `scratch` still holds the prepared activations (F16 tiles of the tensor-core path, or q8 records) of
 * the SAME x from the immediately preceding b200_mul_mat[_ex] call on this scratch, with a weight type of the same preparation class — the
 * q/k/v and gate/up projections of a layer share their input, so only the first of them pays for the conversion.  The caller guarantees
 * that neither x nor scratch changed in between; when the flag cannot be honoured (different path) it is ignored. */
enum { B200_MM_REUSE_ACT = 1 };   /* the scratch still holds the prepared activations of THIS x from the previous call of the same routing class: F16 tiles (more
                                   * than 8 columns) or q8 records of the same group (K-quants: q8_K, q4_0 / q8_0: q8_0; at most 8 columns, no batch dims) */
B200_API int    b200_mul_mat_ex(const b200_tensor * w, const b200_tensor * x, const b200_tensor * dst, void * scratch,
                                size_t scratch_bytes, int flags, void * stream);
/* dst[i] = W[i] . x for 2 or 3 weight matrices over the SAME activations (q / k / v of a layer): one tcgen05 launch over the concatenated m-tiles when every W[i] is
 * a K-quant on the tensor-core path and the merged tile count needs no split-K; otherwise the MUL_MATs run one after the other sharing the activation tiles.
 * (ggml-cuda has no counterpart: ggml_cuda_mul_mat_q is launched once per node, mmq.cu:205.)  scratch >= the largest b200_mul_mat_scratch_bytes of the group. */
B200_API int    b200_mul_mat_multi_merges(int n_mat, const b200_tensor * const * w, const b200_tensor * x);   /* 1: the group runs as one launch */
B200_API int    b200_mul_mat_multi(int n_mat, const b200_tensor * const * w, const b200_tensor * x, const b200_tensor * const * dst, void * scratch,
                                   size_t scratch_bytes, int flags, void * stream);
/* dst = W . x + residual (the ADD behind wo / ffn_down, ggml-cuda fuses nothing here): rides in the tensor-core GEMM's epilogue when it can (quantised weights, 2-D
 * operands, no split-K) and in the decode matvec's epilogue (one activation column, quantised or F16 weights); otherwise MUL_MAT + ADD kernels, which need a
 * residual that does not overlap dst (B200_ERR_UNSUPPORTED before anything is launched).  residual has dst's shape and row stride. */
B200_API int    b200_mul_mat_add(const b200_tensor * w, const b200_tensor * x, const b200_tensor * residual, const b200_tensor * dst, void * scratch,
                                 size_t scratch_bytes, int flags, void * stream);

/* dst = silu(Wg . x) * (Wu . x) for ONE activation column: the gate / up / SWIGLU triple of a decode graph in one launch (quantised or F16 weights; SWIGLU only).
 * B200_ERR_UNSUPPORTED otherwise — the caller keeps MUL_MAT, MUL_MAT, GLU.  scratch >= b200_mul_mat_scratch_bytes(w_gate, x). */
B200_API int    b200_mul_mat_glu(int glu_op, const b200_tensor * w_gate, const b200_tensor * w_up, const b200_tensor * x, const b200_tensor * dst, void * scratch,
                                 size_t scratch_bytes, int flags /* B200_MM_REUSE_ACT: the scratch already holds x's q8 record */, void * stream);

/* Decode fast path: y[m] (+= residual) = W[m, k] . act, activations already quantised by b200_quantize_act /
 * a fused producer.  Up to 4 weight matrices that share the same activation run in ONE launch (q/k/v, gate/up).
 * `residual` may be NULL; when not NULL, y[i] = dot + residual[i] (fuses the ADD that follows wo / ffn_down). */
typedef struct b200_matvec_job {
    const void * w; int32_t type; int32_t layout; int64_t m; int64_t row_stride_bytes;
    float * y; const float * residual;
} b200_matvec_job;
B200_API int b200_matvec_q(const b200_matvec_job * jobs, int njobs, const void * act, int64_t k, void * stream);
/* gate/up pair with the swiglu epilogue fused: y[m] = silu(Wg.act) * (Wu.act)   (GLU: ggml-cuda/unary.cu:208-228) */
B200_API int b200_matvec_q_swiglu(const b200_matvec_job * gate, const b200_matvec_job * up, float * y, const void * act,
                                  int64_t k, void * stream);

/* ---- normalisation (replaces rms_norm_f32<.., do_mul, do_add> ggml-cuda/norm.cu:107-185; oracle ops.cpp:3517-3565) ---
 * y = x * rsqrt(mean(x^2) + eps) [* w] [+ add].  w broadcasts over rows (ne[0] == x->ne[0]) or is NULL. */
B200_API int b200_rms_norm(const b200_tensor * x, const b200_tensor * w, const b200_tensor * add, const b200_tensor * dst,
                           float eps, void * stream);
/* fused producer for the decode path: y = rms_norm(x) * w, quantised straight to act records (one per column; no F32 round
 * trip unless y_f32_or_null is given).  x: ncols columns of k floats, x_col_stride elements apart. */
B200_API int b200_rms_norm_quantize(const float * x, int64_t x_col_stride, const float * w, float * y_f32_or_null,
                                    int64_t y_col_stride, void * act, int weight_type, int64_t k, int64_t ncols, float eps,
                                    void * stream);

/* ---- ROPE (replaces rope_norm / rope_neox ggml-cuda/rope.cu:40-123; oracle ops.cpp:5436-5720) --------------------------- */
typedef struct b200_rope_params {
    int32_t n_dims, mode, n_ctx_orig;      /* mode: 0 = norm, 2 = neox (GGML_ROPE_TYPE_NEOX) */
    float freq_base, freq_scale, ext_factor, attn_factor, beta_fast, beta_slow;
} b200_rope_params;
B200_API int b200_rope(const b200_tensor * x, const int32_t * pos, const float * freq_factors, const b200_tensor * dst,
                       const b200_rope_params * p, void * stream);

/* ---- fused q/k post-processing (replaces, per layer, 2 x rms_norm_f32 norm.cu:107-185 + 2 x rope_neox rope.cu:83-123 +
 *      2 x k_set_rows set-rows.cu:264; graph side: llm_build_qwen3 src/llama-model.cpp:9320-9350, llama-kv-cache.cpp:1021-1110) ---
 * For each of n_tok tokens: q heads are RMS-normed (if q_norm_w), multiplied by the norm weight and rotated IN PLACE; k heads get
 * the same treatment and are written as F16 to k_cache row kv_idx[t]; v is converted to F16 into v_cache row kv_idx[t].
 * Strides: *_tok_stride in elements between tokens; cache row strides in bytes.  head_dim 64 or 128, p->n_dims == head_dim. */
B200_API int b200_qkv_post(float * q, const float * k, const float * v, const float * q_norm_w, const float * k_norm_w,
                           const int32_t * pos, const void * kv_idx, int idx_type, void * k_cache, void * v_cache,
                           int64_t k_row_stride_bytes, int64_t v_row_stride_bytes, int head_dim, int n_head, int n_head_kv,
                           int64_t n_tok, int64_t q_tok_stride, int64_t k_tok_stride, int64_t v_tok_stride,
                           const b200_rope_params * p, float eps, void * stream);

/* ---- KV-cache write / gathers / copies ----------------------------------------------------------------------------------
 * SET_ROWS (set-rows.cu:264; oracle ggml_compute_forward_set_rows): dst[idx[r], :] = convert(src[r, :]), idx I64 or I32.
 * GET_ROWS (getrows.cu:238): dst[r, :] = float(src[idx[r], :]).  CPY (cpy.cu:280): strided copy with F32/F16/BF16 casts. */
B200_API int b200_set_rows(const b200_tensor * src, const b200_tensor * idx, const b200_tensor * dst, void * stream);
B200_API int b200_get_rows(const b200_tensor * src, const b200_tensor * idx, const b200_tensor * dst, void * stream);
B200_API int b200_cpy(const b200_tensor * src, const b200_tensor * dst, void * stream);

/* ---- elementwise (binbcast.cu:395-443, unary.cu, scale.cu) -------------------------------------------------------------- */
enum b200_binop { B200_ADD = 0, B200_SUB = 1, B200_MUL = 2, B200_DIV = 3 };
B200_API int b200_binary(int op, const b200_tensor * a, const b200_tensor * b, const b200_tensor * dst, void * stream);
enum b200_unop { B200_SILU = 0, B200_GELU = 1, B200_RELU = 2, B200_GELU_QUICK = 3, B200_TANH = 4, B200_SIGMOID = 5,
                 B200_GELU_ERF = 6, B200_NEG = 7, B200_EXP = 8, B200_SQR = 9, B200_SQRT = 10, B200_ABS = 11 };
B200_API int b200_unary(int op, const b200_tensor * x, const b200_tensor * dst, void * stream);
/* the rest of unary.cu / clamp.cu the Token2Wav graphs use: SIN / COS / LOG (GGML_OP_SIN ...), ELU / STEP / SGN / HARDSWISH / HARDSIGMOID
 * (GGML_OP_UNARY), LEAKY_RELU (p0 = negative slope; CPU ggml-cpu/ops.cpp:2453-2480), CLAMP (p0 = min, p1 = max; ops.cpp:5309-5340) */
enum b200_unop_ext { B200_SIN = 12, B200_COS = 13, B200_LOG = 14, B200_ELU = 15, B200_STEP = 16, B200_SGN = 17, B200_HARDSWISH = 18, B200_HARDSIGMOID = 19,
                     B200_LEAKY_RELU = 20, B200_CLAMP = 21 };
B200_API int b200_unary_param(int op, const b200_tensor * x, const b200_tensor * dst, float p0, float p1, void * stream);
enum b200_gluop { B200_GLU_REGLU = 0, B200_GLU_GEGLU = 1, B200_GLU_SWIGLU = 2, B200_GLU_GEGLU_ERF = 4, B200_GLU_GEGLU_QUICK = 5 };
/* GLU: dst = act(gate) * up.  up == NULL: single-tensor form, halves of x's rows (swapped selects which half gates). */
B200_API int b200_glu(int op, const b200_tensor * gate_or_x, const b200_tensor * up, const b200_tensor * dst, int swapped,
                      void * stream);
B200_API int b200_scale(const b200_tensor * x, const b200_tensor * dst, float scale, float bias, void * stream);
/* SOFT_MAX (softmax.cu:253): dst = softmax(x*scale + mask) over ne[0]; mask F32/F16 [ne0, ne1, ..] broadcast or NULL */
B200_API int b200_soft_max(const b200_tensor * x, const b200_tensor * mask, const b200_tensor * dst, float scale,
                           float max_bias, void * stream);

/* ---- FLASH_ATTN_EXT (replaces fattn.cu:195-341 -> fattn-vec.cuh / fattn-mma-f16.cuh; oracle ops.cpp:7912-8148) ----------
 * q F32 [D, n_q, n_head, n_b], k/v F16 [D, n_kv, n_head_kv, n_b], mask F16 [n_kv, >= n_q, ...] or NULL,
 * dst F32 [D, n_head, n_q, n_b].  `scratch`: b200_flash_attn_scratch_bytes (split-KV partials / KV-tile counts).
 * n_q < 16: split-KV decode kernel (csrc/flash_attn.cu); n_q >= 16: tiled tensor-core kernel with a mask pre-scan (csrc/fa_prefill.cu,
 * replaces fattn-mma-f16.cuh:1246 + flash_attn_mask_to_KV_max). */
B200_API int    b200_flash_attn_supported(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v,
                                          const b200_tensor * mask, const b200_tensor * dst);
B200_API size_t b200_flash_attn_scratch_bytes(const b200_tensor * q, const b200_tensor * k);
B200_API int    b200_flash_attn(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask,
                                const b200_tensor * dst, float scale, float max_bias, float logit_softcap, void * scratch,
                                size_t scratch_bytes, void * stream);

/* ---- persistent decode engine (replaces, for a batch-1 token, the whole per-node launch sequence ggml_backend_cuda_graph_compute
 *      ggml-cuda.cu:3085-3164 issues for llm_build_qwen3's graph, src/llama-model.cpp:9287-9406: 868 node launches + 253 quantize_q8_1) ---
 * ONE cooperative kernel, one CTA per SM, walks 5 phases per layer + lm_head with grid barriers; weights stream through per-warp TMA
 * rings that keep filling across the barriers (csrc/stream_decode.cu).  All pointers are device pointers and are baked into the program:
 * the host updates the CONTENTS of x_in / pos / kv_idx / mask before each step (what the ggml scheduler's input copies do).
 * Requirements (else B200_ERR_UNSUPPORTED and the caller uses the per-op entry points): head_dim 128, n_head = 4 * n_head_kv, K-quant
 * weights (q4_K / q5_K native, q6_K planar), n_embd and n_ff multiples of 256. */
typedef struct b200_weight { const void * data; int32_t type; int32_t layout; } b200_weight;
typedef struct b200_decode_layer {
    b200_weight wq, wk, wv, wo, gate, up, down;
    const float * attn_norm, * ffn_norm, * q_norm, * k_norm;      /* q_norm/k_norm NULL: no per-head norm (llama arch) */
    void * k_cache, * v_cache;                                   /* F16 [n_ctx][n_head_kv * head_dim] */
    int64_t k_row_bytes, v_row_bytes;
} b200_decode_layer;
typedef struct b200_decode_desc {
    int32_t n_layer, n_embd, n_head, n_head_kv, head_dim, n_ff, n_vocab;
    float rms_eps, attn_scale;
    b200_rope_params rope;
    const b200_decode_layer * layers;
    const float * out_norm; b200_weight lm_head;                 /* lm_head.data NULL: stage without head, x_out receives the residual stream */
    const float * x_in;                                          /* [n_embd] embedding row of the token (or the previous stage's x_out) */
    const int32_t * pos; const int64_t * kv_idx; const void * mask;   /* pos[1], kv_idx[1], F16 mask row [n_kv] (0 / -inf) */
    float * logits; float * hidden_out; float * x_out;           /* [n_vocab]; optional [n_embd] result_norm; optional [n_embd] */
} b200_decode_desc;
B200_API int  b200_decoder_create(const b200_decode_desc * desc, void ** handle);
B200_API int  b200_decoder_step(void * handle, int32_t n_kv, void * stream);
B200_API int  b200_decoder_n_phases(void * handle);
/* debugging aid: one step with per-phase device timestamps of every CTA -> host_out[n_phases][b200_device_sm_count][8]
 * (ns; 0..3: phase start, prologue done, work done, barrier left; 4..7: finer marks) */
B200_API int  b200_decoder_profile(void * handle, int32_t n_kv, unsigned long long * host_out, void * stream);
B200_API void b200_decoder_destroy(void * handle);

/* ---- producers that feed a tensor-core MUL_MAT directly: the [n, k] result is written as that MUL_MAT's prepared F16 activation tiles (into ITS scratch buffer)
 *      instead of (or besides) F32, and the MUL_MAT is then called with B200_MM_REUSE_ACT — no separate conversion pass (k_x_to_f16_tiles), no F32 round trip.
 *      (The reference converts / quantises the activations inside every MUL_MAT: quantize_mmq_q8_1, ggml-cuda/quantize.cu:50-146.)
 * b200_rms_norm_tiles  : RMS_NORM(x) [* w] of a 2-D x [k, n], 1024 < k <= 4096, k % 64 == 0; dst->data may be NULL (tiles only)
 * b200_glu_tiles       : GLU(gate, up) of two contiguous 2-D F32 matrices, k % 64 == 0 (tiles only: the ffn_down MUL_MAT is h's only reader)
 * b200_flash_attn_tiles: FLASH_ATTN_EXT whose [n_q, n_head * 128] result only feeds wo; tcgen05 kernel only, else B200_ERR_UNSUPPORTED; dst->data is not written */
B200_API int b200_rms_norm_tiles(const b200_tensor * x, const b200_tensor * w, const b200_tensor * dst, void * tiles, float eps, void * stream);
B200_API int b200_glu_tiles(int op, const b200_tensor * gate, const b200_tensor * up, void * tiles, void * stream);
B200_API int b200_flash_attn_tiles(const b200_tensor * q, const b200_tensor * k, const b200_tensor * v, const b200_tensor * mask, const b200_tensor * dst,
                                   float scale, void * scratch, size_t scratch_bytes, void * tiles, void * stream);

/* ---- ops of the APM (Whisper) / VPM (SigLip) encoder graphs -----------------------------------------------------------
 * b200_norm   : GGML_OP_NORM, y = (x - mean) / sqrt(var + eps) per row        replaces norm_f32 (ggml-cuda/norm.cu:5); CPU ggml-cpu/ops.cpp:3450-3495
 * b200_im2col : GGML_OP_IM2COL (ggml_conv_1d / ggml_conv_2d), F32 input -> F16 or F32 [IC*KH*KW, OW, OH, N]; `kernel` only supplies KW / KH
 *               replaces im2col_kernel (ggml-cuda/im2col.cu:6); CPU ops.cpp:6160-6301
 * b200_pool_1d: GGML_OP_POOL_1D, op 0 = max, 1 = avg, k0 == s0, p0 == 0 (all the reference implements, ops.cpp:7212-7280; its CUDA backend has none) */
B200_API int b200_norm(const b200_tensor * x, const b200_tensor * dst, float eps, void * stream);
B200_API int b200_im2col(const b200_tensor * kernel, const b200_tensor * x, const b200_tensor * dst, int s0, int s1, int p0, int p1, int d0, int d1,
                         int is_2d, void * stream);
B200_API int b200_pool_1d(const b200_tensor * x, const b200_tensor * dst, int op, int k0, int s0, int p0, void * stream);

/* ---- ops of the Token2Wav graphs (flow-matching CFM + HiFiGAN; csrc/ops_wave.cu).  tools/omni/token2wav/token2wav-impl.cpp:1905-1916 runs
 *      them with a direct graph_compute on the first GPU-type device (no CPU fallback), so the whole op set has to exist ------------------------------------------
 * b200_concat            GGML_OP_CONCAT along dim 0..3, F32 / F16 / BF16 / I32                      replaces ggml-cuda/concat.cu; CPU ggml-cpu/ops.cpp:1839-2040
 * b200_repeat            GGML_OP_REPEAT (dst.ne[i] a multiple of src.ne[i]), same types             replaces binbcast.cu op_repeat; CPU ops.cpp:1637-1700
 * b200_arange            GGML_OP_ARANGE dst[i] = start + step*i (F32, 1-D)                          replaces arange.cu; CPU ops.cpp:7762-7785
 * b200_sum_rows          GGML_OP_SUM_ROWS F32 [ne0, r..] -> [1, r..] (f64 accumulation as the CPU)  replaces sumrows.cu; CPU ops.cpp:1399-1430
 * b200_pad               GGML_OP_PAD zero padding, lp_rp8 = {lp0, rp0, lp1, rp1, lp2, rp2, lp3, rp3} replaces pad.cu; CPU ops.cpp:7592-7640
 * b200_pad_reflect_1d    GGML_OP_PAD_REFLECT_1D                                                     replaces pad_reflect_1d.cu; CPU ops.cpp:7664-7692
 * b200_conv_transpose_1d GGML_OP_CONV_TRANSPOSE_1D (p0 = 0, d0 = 1), kernel [K, Cout, Cin] F32/F16, x [L, Cin] F32 -> [(L-1)*s0 + K, Cout] F32
 *                                                                                                   replaces conv-transpose-1d.cu; CPU ops.cpp:5952-6130 */
B200_API int b200_concat(const b200_tensor * a, const b200_tensor * b, const b200_tensor * dst, int dim, void * stream);
B200_API int b200_repeat(const b200_tensor * src, const b200_tensor * dst, void * stream);
B200_API int b200_arange(const b200_tensor * dst, float start, float step, void * stream);
B200_API int b200_sum_rows(const b200_tensor * x, const b200_tensor * dst, void * stream);
B200_API int b200_pad(const b200_tensor * x, const b200_tensor * dst, const int32_t * lp_rp8, void * stream);
B200_API int b200_pad_reflect_1d(const b200_tensor * x, const b200_tensor * dst, int p0, int p1, void * stream);
B200_API int b200_conv_transpose_1d(const b200_tensor * kernel, const b200_tensor * x, const b200_tensor * dst, int s0, void * stream);

/* ---- layer-split pipeline hop on the device (csrc/hop.cu; replaces, for one-process-per-GPU launches, the per-boundary copy the reference
 *      issues from ggml_backend_sched_compute_splits -> cpy_tensor_async, ggml/src/ggml-backend.cpp:1539, ggml-cuda.cu:2598-2620) -------------
 * b200_ipc_*: device memory another process of the node can map (cudaIpc*); handle64 is the 64-byte cudaIpcMemHandle_t.
 * b200_hop_send: 1-CTA kernel: wait for the consumer's ack of the slot, write n floats into the PEER's input slot, release its ready word.
 * b200_hop_wait: 1-CTA kernel in front of a stage's decode step: acquire the ready word the producer writes.
 * b200_hop_ack : after the step: tell the producer the slot has been consumed.
 * `state` = 8 zero-initialised device bytes per call site (sequence counter + error word: 1 = ack time-out, 2 = ready time-out). */
B200_API int b200_ipc_alloc(size_t bytes, void ** ptr, void * handle64);
B200_API int b200_ipc_open(const void * handle64, void ** ptr);
B200_API int b200_ipc_close(void * ptr);
B200_API int b200_ipc_free(void * ptr);
B200_API int b200_hop_send(const float * src, float * dst_peer, int64_t n, unsigned * ready_peer, const unsigned * ack_local, void * state, void * stream);
B200_API int b200_hop_wait(const unsigned * ready_local, void * state, void * stream);
B200_API int b200_hop_ack(unsigned * ack_peer, void * state, void * stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_OPS_H */
