// llama_parity.cpp — end-to-end parity THROUGH THE REFERENCE: the unmodified libllama (oracle/_ref, built from /root/reference by
// oracle/Makefile) decodes the same synthetic Qwen3 GGUF once on its own ggml CPU backend (n_gpu_layers = 0) and once with every layer
// offloaded to the backend loaded from GGML_BACKEND_PATH (libggml-b200.so), greedy sampling, same prompt.  north_star bar: bit-exact
// token ids for greedy decode, logits within 1e-3 relative.  TEST INFRASTRUCTURE (uses only the reference's public API, include/llama.h).
//
//   llama_parity model.gguf [n_prompt=32] [n_gen=32] [n_threads=8] [flash_attn=1] [mode=0]
//   mode 0: CPU vs B200 (batched prompt)   1: CPU vs CPU-repacked   2: CPU vs CPU 3 threads   3: CPU batched vs CPU incremental
//        4: B200 batched vs B200 incremental   5: B200 per-op route vs B200 decode engine   6: CPU vs B200, prompt fed token by token (decode only)
// prints one JSON line: {"tokens_equal": bool, "n_gen": N, "first_mismatch": i, "max_rel_logit_err": x, "prefill_rel_err": y, ...}
#include "llama.h"
#include <cstring>
#include "ggml-backend.h"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

struct Run { std::vector<llama_token> toks; std::vector<std::vector<float>> logits, hidden; double tg_ms = 0, pp_ms = 0; bool ok = false; };

static Run run(const char * path, int ngl, int n_prompt, int n_gen, int n_threads, int fa, const std::vector<llama_token> * force, bool repack, bool incremental = false) {
    Run r;
    llama_model_params mp = llama_model_default_params();
    static ggml_backend_dev_t no_devices[1] = { nullptr };
    if (ngl == 0) mp.devices = no_devices;         // a true CPU run: otherwise batches >= 32 tokens are offloaded to the GPU backend op by op (offload_op)
    mp.n_gpu_layers = ngl; mp.use_mmap = true; mp.use_extra_bufts = repack;      // repack = the CPU backend's interleaved q4_K_8x8 weights (different f32 summation order)
    llama_model * model = llama_model_load_from_file(path, mp);
    if (!model) { fprintf(stderr, "load failed\n"); return r; }
    llama_context_params cp = llama_context_default_params();
    const int nb = n_prompt > 512 ? (n_prompt + 255) / 256 * 256 : 512;                 // the whole prompt as ONE ubatch (exercises the n >= 512 GEMM routing)
    cp.n_ctx = n_prompt + n_gen + 64 > 1024 ? (n_prompt + n_gen + 64 + 255) / 256 * 256 : 1024; cp.n_batch = nb; cp.n_ubatch = nb; cp.n_threads = n_threads; cp.n_threads_batch = n_threads;
    cp.flash_attn_type = fa ? LLAMA_FLASH_ATTN_TYPE_ENABLED : LLAMA_FLASH_ATTN_TYPE_DISABLED; cp.no_perf = true;
    // PARITY_KV_TYPE=q8_0|q4_0: quantised KV cache on both sides (-ctk / -ctv: SET_ROWS into quant blocks, FLASH_ATTN_EXT over them; needs fa = 1)
    if (const char * kvt = getenv("PARITY_KV_TYPE")) {
        const ggml_type t = strcmp(kvt, "q8_0") == 0 ? GGML_TYPE_Q8_0 : strcmp(kvt, "q4_0") == 0 ? GGML_TYPE_Q4_0 : GGML_TYPE_F16;
        cp.type_k = t; cp.type_v = t;
    }
    llama_context * ctx = llama_init_from_model(model, cp);
    if (!ctx) { fprintf(stderr, "context failed\n"); llama_model_free(model); return r; }
    const int n_vocab = llama_vocab_n_tokens(llama_model_get_vocab(model));
    // PARITY_EMBEDDINGS=1: omni's stream_decode pattern (tools/omni/omni.cpp:889-916, eval_id_with_hidden): llama_set_embeddings(ctx, true) around every llama_decode, so
    // the graph also outputs result_norm (src/llama-model.cpp:9395-9396) and the caller reads the hidden state next to the logits (the TTS input)
    const bool want_hidden = getenv("PARITY_EMBEDDINGS") && atoi(getenv("PARITY_EMBEDDINGS")) != 0;
    const int n_embd = llama_model_n_embd(model);
    auto decode = [&](llama_batch b) {
        if (want_hidden) llama_set_embeddings(ctx, true);
        const int rc = llama_decode(ctx, b);
        if (want_hidden) {
            if (rc == 0) { const float * e = llama_get_embeddings_ith(ctx, -1); if (e) r.hidden.emplace_back(e, e + n_embd); }
            llama_set_embeddings(ctx, false);
        }
        return rc;
    };
    std::vector<llama_token> prompt(n_prompt);
    uint32_t s = 12345;
    for (auto & t : prompt) { s = s * 1664525u + 1013904223u; t = (llama_token) ((s >> 8) % (uint32_t) n_vocab); }
    auto t0 = std::chrono::steady_clock::now();
    if (incremental) {                                                  // the same prompt one token at a time (n = 1 graphs only)
        for (int i = 0; i < n_prompt; ++i) if (decode(llama_batch_get_one(&prompt[i], 1))) { fprintf(stderr, "prefill failed\n"); return r; }
    } else if (decode(llama_batch_get_one(prompt.data(), n_prompt))) { fprintf(stderr, "prefill failed\n"); return r; }
    const float * lg = llama_get_logits_ith(ctx, -1);
    r.logits.emplace_back(lg, lg + n_vocab);
    // PARITY_KSHIFT=1: omni's sliding window (tools/omni/omni.cpp:686-820): drop a quarter of the prompt behind the first token and shift the rest down; the next
    // llama_decode then runs the K-shift graph (ROPE in place on views of the cache: F16 directly, a quantised cache through F32 casts, src/llama-kv-cache.cpp)
    if (getenv("PARITY_KSHIFT") && atoi(getenv("PARITY_KSHIFT")) != 0 && n_prompt >= 8) {
        llama_memory_t mem = llama_get_memory(ctx);
        const int n_discard = n_prompt / 4;
        llama_memory_seq_rm(mem, 0, 1, 1 + n_discard);
        llama_memory_seq_add(mem, 0, 1 + n_discard, n_prompt, -n_discard);
    }
    auto t1 = std::chrono::steady_clock::now();
    r.pp_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    for (int i = 0; i < n_gen; ++i) {
        const std::vector<float> & l = r.logits.back();
        llama_token best = 0;
        for (int v = 1; v < n_vocab; ++v) if (l[v] > l[best]) best = v;
        r.toks.push_back(best);
        // teacher forcing on the second run keeps the two runs on the same sequence even after a (reported) mismatch
        llama_token feed = force && i < (int) force->size() ? (*force)[i] : best;
        if (decode(llama_batch_get_one(&feed, 1))) { fprintf(stderr, "decode failed\n"); return r; }
        lg = llama_get_logits_ith(ctx, -1);
        r.logits.emplace_back(lg, lg + n_vocab);
    }
    r.tg_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    r.ok = true;
    llama_free(ctx); llama_model_free(model);
    return r;
}

int main(int argc, char ** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s model.gguf [n_prompt] [n_gen] [threads] [fa]\n", argv[0]); return 2; }
    const int n_prompt = argc > 2 ? atoi(argv[2]) : 32, n_gen = argc > 3 ? atoi(argv[3]) : 32, nt = argc > 4 ? atoi(argv[4]) : 8, fa = argc > 5 ? atoi(argv[5]) : 1;
    llama_log_set([](ggml_log_level lvl, const char * txt, void *) { if (lvl >= GGML_LOG_LEVEL_ERROR) fputs(txt, stderr); }, nullptr);
    ggml_backend_load_all();
    llama_backend_init();
    size_t n_gpu = 0;
    for (size_t i = 0; i < ggml_backend_dev_count(); ++i) if (ggml_backend_dev_type(ggml_backend_dev_get(i)) == GGML_BACKEND_DEVICE_TYPE_GPU) ++n_gpu;
    // mode "cpu-self": the CPU backend against itself (plain vs repacked weights) — how far two correct implementations drift on this model
    const bool cpu_self = argc > 6 && atoi(argv[6]) == 1;
    if (!n_gpu && !cpu_self && !(argc > 6 && (atoi(argv[6]) == 2 || atoi(argv[6]) == 3))) { printf("{\"error\": \"no GPU backend registered (GGML_BACKEND_PATH?)\"}\n"); return 3; }
    const int self_mode = argc > 6 ? atoi(argv[6]) : 0;
    // 6: DECODE-ONLY parity — the prompt is fed one token at a time on BOTH sides, so only n = 1 graphs (the decode path, which carries the
    //    reference's integer arithmetic) ever run; the batched-prefill kernels (F16 operands on the GPU vs q8_K on the CPU) are out of the picture
    Run cpu = run(argv[1], 0, n_prompt, n_gen, nt, fa, nullptr, false, self_mode == 6);                 // 1: repacked weights; 2: plain weights, 3 threads (must be bit-identical)
    // 3: CPU batched prompt vs CPU one-token-at-a-time prompt; 4: B200 batched vs B200 one-at-a-time (baseline run also on the GPU)
    if (self_mode == 4) cpu = run(argv[1], 999, n_prompt, n_gen, nt, fa, nullptr, false);
    // 7: RE-ENTRANCY — three host threads, each with its OWN model + context (three backend instances, three streams, three decode engines with their
    //    cooperative 148-CTA launches) decode the same sequence concurrently on one device; every thread must reproduce the single-threaded B200 run
    //    BIT FOR BIT (same kernels, deterministic reductions).  This is what omni_init's LLM / TTS / encoder threads do to a backend (omni.h:287).
    if (self_mode == 7) {
        Run base = run(argv[1], 999, n_prompt, n_gen, nt, fa, nullptr, false);
        Run th[3];
        std::thread ts[3];
        for (int i = 0; i < 3; ++i) ts[i] = std::thread([&, i] { th[i] = run(argv[1], 999, n_prompt, n_gen, nt, fa, nullptr, false, i == 2); });
        for (auto & t : ts) t.join();
        bool ok = base.ok; double worst = 0; int bad = 0;
        for (int i = 0; i < 3 && ok; ++i) {
            ok = ok && th[i].ok;
            if (!ok) break;
            const bool incr = i == 2;                                  // thread 2 feeds the prompt token by token (different graphs in flight at the same time)
            for (size_t s2 = incr ? 1 : 0; s2 < base.logits.size(); ++s2)
                for (size_t v = 0; v < base.logits[s2].size(); ++v) {
                    const double e = std::fabs(base.logits[s2][v] - th[i].logits[s2][v]);
                    if (!incr && e != 0) ++bad;
                    worst = std::fmax(worst, e);
                }
            if (!incr && th[i].toks != base.toks) ++bad;
        }
        printf("{\"mode\": 7, \"threads_ok\": %s, \"bitwise_mismatches\": %d, \"max_abs_diff_incl_incremental_thread\": %.3g}\n", ok ? "true" : "false", bad, worst);
        return ok && bad == 0 ? 0 : 1;
    }
    // 8: B200 with the producer -> MUL_MAT tile fusion of n-token graphs (RMS_NORM / GLU / FLASH_ATTN_EXT write the next MUL_MAT's F16 tiles) vs without it:
    //    the fused tiles hold the same F16 roundings of the same F32 values, so the logits must be BIT-IDENTICAL
    if (self_mode == 8) { setenv("GGML_B200_NO_TILE_FUSION", "1", 1); cpu = run(argv[1], 999, n_prompt, n_gen, nt, fa, nullptr, false); setenv("GGML_B200_NO_TILE_FUSION", "0", 1); }
    // 5: B200 per-op launches (baseline) vs B200 whole-token decode engine — the plugin reads GGML_B200_DISABLE_ENGINE when a backend is created
    if (self_mode == 5) { setenv("GGML_B200_DISABLE_ENGINE", "1", 1); cpu = run(argv[1], 999, n_prompt, n_gen, nt, fa, nullptr, false); setenv("GGML_B200_DISABLE_ENGINE", "0", 1); }
    Run gpu = run(argv[1], cpu_self || self_mode == 2 || self_mode == 3 ? 0 : 999, n_prompt, n_gen, self_mode == 2 ? 3 : nt, fa, &cpu.toks, cpu_self, self_mode == 3 || self_mode == 4 || self_mode == 6);
    if (!cpu.ok || !gpu.ok) { printf("{\"error\": \"run failed\"}\n"); return 4; }
    int first = -1; double max_rel = 0, pre_rel = 0; double step_rel[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n_gen; ++i) if (cpu.toks[i] != gpu.toks[i] && first < 0) first = i;
    for (size_t i = 0; i < cpu.logits.size(); ++i) {
        double mx = 0, err = 0;
        for (size_t v = 0; v < cpu.logits[i].size(); ++v) { mx = std::fmax(mx, std::fabs(cpu.logits[i][v])); err = std::fmax(err, std::fabs(cpu.logits[i][v] - gpu.logits[i][v])); }
        const double rel = err / (mx > 0 ? mx : 1);
        if (i == 0) pre_rel = rel; else max_rel = std::fmax(max_rel, rel);
        if (i == 0 && self_mode == 6) max_rel = std::fmax(max_rel, rel);   // decode-only mode: the prompt's last logits come from n = 1 graphs too
        if (i >= 1 && i <= 6) step_rel[i - 1] = rel;
    }
    // hidden states (PARITY_EMBEDDINGS=1): one per llama_decode on both sides, compared relative to each vector's largest magnitude
    double hid_rel = -1;
    if (!cpu.hidden.empty() && cpu.hidden.size() == gpu.hidden.size()) {
        hid_rel = 0;
        for (size_t i = 0; i < cpu.hidden.size(); ++i) {
            double mx = 0, err = 0;
            for (size_t v = 0; v < cpu.hidden[i].size(); ++v) { mx = std::fmax(mx, std::fabs(cpu.hidden[i][v])); err = std::fmax(err, std::fabs(cpu.hidden[i][v] - gpu.hidden[i][v])); }
            hid_rel = std::fmax(hid_rel, err / (mx > 0 ? mx : 1));
        }
        printf("{\"hidden_states\": %zu, \"max_rel_hidden_err\": %.3g}\n", cpu.hidden.size(), hid_rel);
    } else if (cpu.hidden.size() != gpu.hidden.size()) { printf("{\"error\": \"hidden state count differs: %zu vs %zu\"}\n", cpu.hidden.size(), gpu.hidden.size()); return 4; }
    printf("{\"tokens_equal\": %s, \"n_prompt\": %d, \"n_gen\": %d, \"first_mismatch\": %d, \"max_rel_logit_err\": %.3g, \"prefill_rel_err\": %.3g, "
           "\"decode_step_rel_err\": [%.2g, %.2g, %.2g, %.2g, %.2g, %.2g], \"cpu_tg_tok_s\": %.2f, \"gpu_tg_tok_s\": %.2f, \"cpu_pp_tok_s\": %.1f, \"gpu_pp_tok_s\": %.1f, \"threads\": %d, \"flash_attn\": %d, \"mode\": %d}\n",
           first < 0 ? "true" : "false", n_prompt, n_gen, first, max_rel, pre_rel, step_rel[0], step_rel[1], step_rel[2], step_rel[3], step_rel[4], step_rel[5], n_gen * 1e3 / cpu.tg_ms, n_gen * 1e3 / gpu.tg_ms,
           n_prompt * 1e3 / cpu.pp_ms, n_prompt * 1e3 / gpu.pp_ms, nt, fa, self_mode);
    return first < 0 && max_rel <= 1e-3 && hid_rel <= 1e-3 ? 0 : 1;
}
