// decoder.cu — C-ABI of the persistent decode engine (stream_decode.cu): builds the phase program of a whole batch-1 decode step
// (5 phases per transformer layer + the lm_head phase) from plain weight / cache pointers and launches it as ONE cooperative kernel.
//
// Graph shape mirrored: llm_build_qwen3 (src/llama-model.cpp:9287-9406) / llm_build_llama without biases:
//   A  RMS_NORM*attn_norm -> q8_K            | wq, wk, wv                                  (3 matrices, one row space)
//   B  q/k RMS_NORM*w, RoPE, KV write, FLASH_ATTN_EXT (split-KV + merge)                   (stream_attn.cuh)
//   C  q8_K(attn)                            | wo + residual
//   D  RMS_NORM*ffn_norm -> q8_K             | gate, up -> SWIGLU
//   E  q8_K(h[K-slice])                      | down, K-split into partials (summed by the next phase's prologue together with the residual)
//   Z  sum -> RMS_NORM*output_norm -> q8_K   | lm_head
#include "stream_decode.cuh"
#include <math.h>
#include <vector>
#include <new>

namespace b200 {
bool sd_fill_mat(SdMat & M, const void * w, int type, int layout, int64_t m, int64_t k, int64_t row_stride_bytes, float * y, const float * residual);
bool sd_phase_ok(const SdPhase & P);
int  sd_launch(const SdPhase * phases_dev, int n_phases, const SdPhase * single, unsigned * gbar, const SdRuntime & rt, cudaStream_t st);

struct Decoder {
    int n_phases = 0;
    SdPhase * phases_dev = nullptr;
    char * scratch = nullptr;
    unsigned * gbar = nullptr;
    SdRuntime rt = {};
    unsigned long long * prof_dev = nullptr;
    ~Decoder() { if (phases_dev) cudaFree(phases_dev); if (scratch) cudaFree(scratch); if (prof_dev) cudaFree(prof_dev); }
};

static int64_t row_bytes(int type, int64_t k) { return k / blck_size(type) * type_size(type); }
// activation record a weight type multiplies with: q8_K per 256 for the K-quants, q8_0 per 32 for q4_0 / q8_0 (the CPU backend's vec_dot_type, ggml-cpu.c:196-350);
// every matrix of one phase must want the same record
static int act_group_of(int type) { return is_kquant(type) ? 256 : 32; }
static float yarn_dim(int n_dims, int n_ctx_orig, float n_rot, float base) { return n_dims * logf(n_ctx_orig / (n_rot * 2 * (float) M_PI)) / (2 * logf(base)); }
} // namespace b200

using namespace b200;

extern "C" int b200_decoder_create(const b200_decode_desc * d, void ** handle) {
    if (!d || !handle || !d->layers || d->n_layer < 0 || !d->x_in || !d->pos || !d->kv_idx || !d->mask) return B200_ERR_ARG;
    const int E = d->n_embd, F = d->n_ff, D = d->head_dim, H = d->n_head, HK = d->n_head_kv, Q = H * D, KV = HK * D;
    if (D != 128 || HK <= 0 || H != HK * 4 || E % 256 || F % 256 || Q % 256) return B200_ERR_UNSUPPORTED;
    if (d->rope.n_dims != D || (d->rope.mode != 0 && d->rope.mode != 2)) return B200_ERR_UNSUPPORTED;
    int nsm = sm_count();
    if (HK > nsm) return B200_ERR_UNSUPPORTED;
    // K-split of ffn_down: slices of at most 4096 so that the activation fragments stay in registers
    int ks = 1;
    while (F / ks > 4096 && ks < 8) ++ks;
    while (ks < 8 && (F % (ks * 256) || F / ks > 4096)) ++ks;
    if (F % (ks * 256) || F / ks > 4096) ks = 1;

    // ---- scratch: activations between phases, K-split partials, attention partials, tickets, grid barrier -------------------------------
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t) 255; return o; };
    const size_t o_q = take(Q * 4), o_k = take(KV * 4), o_v = take(KV * 4), o_attn = take(Q * 4), o_h = take((size_t) F * 4);
    const size_t o_x1 = take(E * 4), o_x2 = take(E * 4), o_xres = take(E * 4), o_part = take((size_t) 8 * E * 4);
    const int max_splits = nsm / HK;
    const size_t o_pacc = take((size_t) H * max_splits * D * 4), o_pml = take((size_t) H * max_splits * 8), o_tick = take(HK * 4 + 256), o_bar = take(512);
    Decoder * dec = new (std::nothrow) Decoder();
    if (!dec) return B200_ERR_ARG;
    cudaError_t e = cudaMalloc(&dec->scratch, off);
    if (e != cudaSuccess) { delete dec; return -(int) e; }
    cudaMemset(dec->scratch, 0, off);
    char * S = dec->scratch;
    float * q = (float *) (S + o_q), * k = (float *) (S + o_k), * v = (float *) (S + o_v), * attn = (float *) (S + o_attn), * h = (float *) (S + o_h);
    float * x1 = (float *) (S + o_x1), * x2 = (float *) (S + o_x2), * xres = (float *) (S + o_xres), * part = (float *) (S + o_part);
    dec->gbar = (unsigned *) (S + o_bar);

    const float lo = floorf(yarn_dim(D, d->rope.n_ctx_orig, d->rope.beta_fast, d->rope.freq_base));
    const float hi = ceilf (yarn_dim(D, d->rope.n_ctx_orig, d->rope.beta_slow, d->rope.freq_base));

    std::vector<SdPhase> ph;
    bool ok = true;
    // the residual stream entering a layer: one vector, or x1 + K-split partials of the previous ffn_down
    const float * xin[4] = { d->x_in, nullptr, nullptr, nullptr }; int n_x = 1;
    auto set_x = [&](SdPhase & P) { for (int i = 0; i < 4; ++i) P.x[i] = xin[i]; P.n_x = n_x; };
    for (int il = 0; il < d->n_layer && ok; ++il) {
        const b200_decode_layer & L = d->layers[il];
        const float * resid_in = n_x == 1 ? xin[0] : xres;               // what wo adds back: the layer input
        {   // A
            SdPhase P = {}; P.kind = SD_MATVEC; P.n_mat = 3; P.epilogue = SD_EPI_STORE; P.prologue = SD_PRO_RMSNORM_QUANT; P.k = E; P.act_group = act_group_of(L.wq.type);
            P.eps = d->rms_eps; P.ksplit = 1; set_x(P); P.x_out = n_x == 1 ? nullptr : xres; P.norm_w = L.attn_norm;
            ok = ok && L.attn_norm && is_quant(L.wq.type) && is_quant(L.wk.type) && is_quant(L.wv.type) && act_group_of(L.wk.type) == P.act_group && act_group_of(L.wv.type) == P.act_group;
            ok = ok && sd_fill_mat(P.mat[0], L.wq.data, L.wq.type, L.wq.layout, Q, E, row_bytes(L.wq.type, E), q, nullptr);
            ok = ok && sd_fill_mat(P.mat[1], L.wk.data, L.wk.type, L.wk.layout, KV, E, row_bytes(L.wk.type, E), k, nullptr);
            ok = ok && sd_fill_mat(P.mat[2], L.wv.data, L.wv.type, L.wv.layout, KV, E, row_bytes(L.wv.type, E), v, nullptr);
            ok = ok && sd_phase_ok(P); ph.push_back(P);
        }
        {   // B
            SdPhase P = {}; P.kind = SD_ATTN; P.ksplit = 1;
            SdAttn & A = P.attn;
            A.q = q; A.k_new = k; A.v_new = v; A.q_norm_w = L.q_norm; A.k_norm_w = L.k_norm;
            A.k_cache = (uint8_t *) L.k_cache; A.v_cache = (uint8_t *) L.v_cache; A.k_row_bytes = L.k_row_bytes; A.v_row_bytes = L.v_row_bytes;
            A.out = attn; A.part_acc = (float *) (S + o_pacc); A.part_ml = (float2 *) (S + o_pml); A.tickets = (unsigned *) (S + o_tick);
            A.n_head = H; A.n_head_kv = HK; A.head_dim = D; A.rope_mode = d->rope.mode; A.scale = d->attn_scale; A.eps = d->rms_eps;
            A.theta_scale = powf(d->rope.freq_base, -2.0f / D); A.freq_scale = d->rope.freq_scale; A.ext_factor = d->rope.ext_factor;
            A.attn_factor = d->rope.attn_factor; A.corr0 = lo < 0 ? 0 : lo; A.corr1 = hi > D - 1 ? D - 1 : hi;
            ok = ok && L.k_cache && L.v_cache && (L.k_row_bytes % 16 == 0) && (L.v_row_bytes % 16 == 0) && ((L.q_norm == nullptr) == (L.k_norm == nullptr));
            ok = ok && ((uintptr_t) L.k_cache % 16 == 0) && ((uintptr_t) L.v_cache % 16 == 0);
            ph.push_back(P);
        }
        {   // C
            SdPhase P = {}; P.kind = SD_MATVEC; P.n_mat = 1; P.epilogue = SD_EPI_STORE; P.prologue = SD_PRO_QUANT; P.k = Q; P.act_group = act_group_of(L.wo.type); P.ksplit = 1;
            P.x[0] = attn; P.n_x = 1;
            ok = ok && is_quant(L.wo.type) && sd_fill_mat(P.mat[0], L.wo.data, L.wo.type, L.wo.layout, E, Q, row_bytes(L.wo.type, Q), x1, resid_in);
            ok = ok && sd_phase_ok(P); ph.push_back(P);
        }
        {   // D
            SdPhase P = {}; P.kind = SD_MATVEC; P.n_mat = 2; P.epilogue = SD_EPI_SWIGLU; P.prologue = SD_PRO_RMSNORM_QUANT; P.k = E; P.act_group = act_group_of(L.gate.type);
            P.eps = d->rms_eps; P.ksplit = 1; P.x[0] = x1; P.n_x = 1; P.norm_w = L.ffn_norm;
            ok = ok && L.ffn_norm && is_quant(L.gate.type) && L.gate.type == L.up.type;
            ok = ok && sd_fill_mat(P.mat[0], L.gate.data, L.gate.type, L.gate.layout, F, E, row_bytes(L.gate.type, E), h, nullptr);
            ok = ok && sd_fill_mat(P.mat[1], L.up.data, L.up.type, L.up.layout, F, E, row_bytes(L.up.type, E), h, nullptr);
            ok = ok && sd_phase_ok(P); ph.push_back(P);
        }
        {   // E
            SdPhase P = {}; P.kind = SD_MATVEC; P.n_mat = 1; P.epilogue = SD_EPI_STORE; P.prologue = SD_PRO_QUANT; P.k = F; P.act_group = act_group_of(L.down.type); P.ksplit = ks;
            P.x[0] = h; P.n_x = 1; P.y_part_stride = ks > 1 ? E : 0;
            ok = ok && is_quant(L.down.type);
            ok = ok && sd_fill_mat(P.mat[0], L.down.data, L.down.type, L.down.layout, E, F, row_bytes(L.down.type, F), ks > 1 ? part : x2, ks > 1 ? nullptr : x1);
            ok = ok && sd_phase_ok(P); ph.push_back(P);
        }
        if (ks > 1) { xin[0] = x1; for (int s = 0; s < 3; ++s) xin[1 + s] = nullptr; n_x = 1 + ks; if (n_x > 4) ok = false; for (int s = 0; s < ks && s < 3; ++s) xin[1 + s] = part + (size_t) s * E; }
        else        { xin[0] = x2; n_x = 1; }
    }
    if (ok) {   // Z: lm_head (or, for a pipeline stage without head, just materialise the summed residual stream)
        SdPhase P = {}; P.kind = SD_MATVEC; P.n_mat = d->lm_head.data ? 1 : 0; P.epilogue = SD_EPI_STORE; P.k = E; P.act_group = d->lm_head.data ? act_group_of(d->lm_head.type) : 256; P.ksplit = 1;
        P.prologue = d->lm_head.data ? SD_PRO_RMSNORM_QUANT : SD_PRO_QUANT; P.eps = d->rms_eps; set_x(P); P.x_out = d->x_out; P.norm_w = d->out_norm; P.norm_out = d->hidden_out;
        if (d->lm_head.data) {
            ok = ok && d->out_norm && d->logits && is_quant(d->lm_head.type);
            ok = ok && sd_fill_mat(P.mat[0], d->lm_head.data, d->lm_head.type, d->lm_head.layout, d->n_vocab, E, row_bytes(d->lm_head.type, E), d->logits, nullptr);
        } else ok = ok && d->x_out;
        ok = ok && sd_phase_ok(P); ph.push_back(P);
    }
    if (!ok) { delete dec; return B200_ERR_UNSUPPORTED; }
    for (size_t p = 0; p < ph.size(); ++p) {                             // let every phase know what follows it (no descriptor reads on the critical path)
        ph[p].next_kind = p + 1 < ph.size() ? ph[p + 1].kind : -1;
        ph[p].next_mv = -1;
        for (size_t n = p + 1; n < ph.size(); ++n) if (ph[n].kind == SD_MATVEC) { ph[p].next_mv = (int) n; break; }
    }
    dec->n_phases = (int) ph.size();
    e = cudaMalloc(&dec->phases_dev, ph.size() * sizeof(SdPhase));
    if (e == cudaSuccess) e = cudaMemcpy(dec->phases_dev, ph.data(), ph.size() * sizeof(SdPhase), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete dec; return -(int) e; }
    dec->rt.pos = d->pos; dec->rt.kv_idx = d->kv_idx; dec->rt.mask = (const __half *) d->mask;
    dec->rt.theta_scale = powf(d->rope.freq_base, -2.0f / D); dec->rt.freq_scale = d->rope.freq_scale; dec->rt.ext_factor = d->rope.ext_factor;
    dec->rt.attn_factor = d->rope.attn_factor; dec->rt.corr0 = lo < 0 ? 0 : lo; dec->rt.corr1 = hi > D - 1 ? D - 1 : hi;
    dec->rt.rope_mode = d->rope.mode; dec->rt.has_rope = d->n_layer > 0;
    // the memset / upload above ran on the legacy stream; k_stream is launched on the caller's (possibly non-blocking) stream, which has no
    // implicit ordering with it: the barrier counters and the phase table must have landed before this function returns
    e = cudaStreamSynchronize(nullptr);
    if (e != cudaSuccess) { delete dec; return -(int) e; }
    *handle = dec;
    return B200_OK;
}

extern "C" int b200_decoder_step(void * handle, int32_t n_kv, void * stream) {
    Decoder * dec = (Decoder *) handle;
    if (!dec || n_kv <= 0) return B200_ERR_ARG;
    SdRuntime rt = dec->rt; rt.n_kv = n_kv;
    return sd_launch(dec->phases_dev, dec->n_phases, nullptr, dec->gbar, rt, (cudaStream_t) stream);
}

// debugging aid: run one step with per-phase globaltimer stamps of every CTA ([n_phases][n_sm][8] ns: start, prologue done, work done, barrier done, 4 finer marks)
extern "C" int b200_decoder_profile(void * handle, int32_t n_kv, unsigned long long * host_out, void * stream) {
    Decoder * dec = (Decoder *) handle;
    if (!dec || !host_out) return B200_ERR_ARG;
    const size_t bytes = (size_t) dec->n_phases * sm_count() * 8 * sizeof(unsigned long long);
    if (!dec->prof_dev) B200_CUDA_TRY(cudaMalloc(&dec->prof_dev, bytes));
    B200_CUDA_TRY(cudaMemsetAsync(dec->prof_dev, 0, bytes, (cudaStream_t) stream));
    SdRuntime rt = dec->rt; rt.n_kv = n_kv; rt.prof = dec->prof_dev;
    int rc = sd_launch(dec->phases_dev, dec->n_phases, nullptr, dec->gbar, rt, (cudaStream_t) stream);
    if (rc) return rc;
    B200_CUDA_TRY(cudaMemcpyAsync(host_out, dec->prof_dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t) stream));
    B200_CUDA_TRY(cudaStreamSynchronize((cudaStream_t) stream));
    return B200_OK;
}

extern "C" int b200_decoder_n_phases(void * handle) { return handle ? ((Decoder *) handle)->n_phases : 0; }
extern "C" void b200_decoder_destroy(void * handle) { delete (Decoder *) handle; }
