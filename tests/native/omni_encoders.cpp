// tests/native/omni_encoders.cpp — TEST INFRASTRUCTURE.  BASELINE.json configs[3] ("APM 1 s audio chunk + VPM frame" in front of the LLM):
// drives the reference's UNMODIFIED omni encoders (tools/omni/audition.cpp, tools/omni/vision.cpp, compiled from where they lie by oracle/Makefile) through their own
// public API on a synthetic GGUF (tools/make_omni_gguf.py), once with use_gpu = false (the reference CPU backend) and once with use_gpu = true
// (ggml_backend_init_by_type(GPU) -> whichever GPU-type backend is registered: libggml-b200.so through GGML_BACKEND_PATH or the LD_PRELOAD shim).  Each encoder owns a
// ggml_backend_sched(GPU, CPU) (audition.cpp:236-266, vision.cpp:196-227), so nodes our supports_op refuses stay on the CPU: GGML_SCHED_DEBUG=2 prints the splits.
//
//   omni_encoders apm <apm.gguf> <n_chunks> [n_threads]     n_chunks successive 1 s chunks (100 mel frames x 80 bins -> 50 tokens -> 10 pooled embeddings each), streaming:
//                                                           chunk i attends to the encoder KV cache of chunks 0..i (audition.cpp:475-640)
//   omni_encoders vpm <vpm.gguf> <n_frames> [n_threads] [w h]   n_frames images of w x h pixels (default 448 x 448 = 32 x 32 patches of 14) -> 64 resampler queries each
//
// prints ONE JSON line: per-chunk milliseconds of both sides and the largest relative difference of the embeddings (max |gpu - cpu| / max |cpu|).
// OMNI_NO_LOAD_ALL=1 skips ggml_backend_load_all() — what llama-omni-cli does (tools/omni/omni-cli.cpp never calls it): the backend must then come from the preload shim.
#include "ggml.h"
#include "ggml-backend.h"
#include "audition.h"
#include "vision.h"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Lcg {
    uint32_t s;
    float next() { s = s * 1664525u + 1013904223u; return (float) (s >> 8) / 8388608.0f - 1.0f; }      // [-1, 1)
};

struct Cmp { double max_ref = 0, max_diff = 0, sse = 0, ssr = 0; size_t n_bad = 0; };
static void compare(Cmp & c, const std::vector<float> & ref, const std::vector<float> & got) {
    for (size_t i = 0; i < ref.size(); ++i) {
        if (!std::isfinite(got[i]) || !std::isfinite(ref[i])) { ++c.n_bad; continue; }
        const double d = (double) got[i] - ref[i];
        c.max_ref = fmax(c.max_ref, fabs(ref[i])); c.max_diff = fmax(c.max_diff, fabs(d)); c.sse += d * d; c.ssr += (double) ref[i] * ref[i];
    }
}

static std::string join(const std::vector<double> & v) {
    std::string s = "[";
    char b[32];
    for (size_t i = 0; i < v.size(); ++i) { snprintf(b, sizeof b, "%s%.3f", i ? ", " : "", v[i]); s += b; }
    return s + "]";
}

static int run_apm(const char * fname, int n_chunks, int n_threads) {
    audition_ctx * cpu = audition_init(fname, { /*use_gpu*/ false, GGML_LOG_LEVEL_ERROR });
    audition_ctx * gpu = audition_init(fname, { /*use_gpu*/ true,  GGML_LOG_LEVEL_ERROR });
    if (!cpu || !gpu) { printf("{\"error\": \"audition_init failed\"}\n"); return 1; }
    const int n_frames = 100, n_mel = 80, n_embd = audition_n_mmproj_embd(cpu);
    Lcg rng{ 4242 };
    Cmp cmp;
    std::vector<double> ms_cpu, ms_gpu;
    int n_tok = 0;
    for (int c = 0; c < n_chunks; ++c) {
        audition_audio_f32 mel;
        mel.nx = n_frames; mel.ny = n_mel; mel.buf.resize((size_t) n_frames * n_mel);
        for (auto & v : mel.buf) v = rng.next();                                           // Whisper's log-mel is normalised to about [-1, 1.5]
        n_tok = audition_n_output_tokens(cpu, &mel);
        std::vector<float> out_cpu((size_t) n_tok * n_embd), out_gpu(out_cpu.size());
        double t0 = now_ms();
        if (!audition_audio_encode(cpu, n_threads, &mel, out_cpu.data())) { printf("{\"error\": \"cpu encode failed at chunk %d\"}\n", c); return 1; }
        double t1 = now_ms();
        if (!audition_audio_encode(gpu, n_threads, &mel, out_gpu.data())) { printf("{\"error\": \"gpu encode failed at chunk %d\"}\n", c); return 1; }
        double t2 = now_ms();
        ms_cpu.push_back(t1 - t0); ms_gpu.push_back(t2 - t1);
        compare(cmp, out_cpu, out_gpu);
    }
    printf("{\"encoder\": \"apm\", \"chunks\": %d, \"tokens_per_chunk\": %d, \"n_embd\": %d, \"threads\": %d, \"max_rel_err\": %.3e, \"nmse\": %.3e, \"non_finite\": %zu, "
           "\"ms_cpu\": %s, \"ms_gpu\": %s}\n", n_chunks, n_tok, n_embd, n_threads, cmp.max_diff / fmax(cmp.max_ref, 1e-30), cmp.sse / fmax(cmp.ssr, 1e-30), cmp.n_bad,
           join(ms_cpu).c_str(), join(ms_gpu).c_str());
    audition_free(gpu); audition_free(cpu);
    return 0;
}

static int run_vpm(const char * fname, int n_frames, int n_threads, int w, int h) {
    vision_ctx * cpu = vision_init(fname, { /*use_gpu*/ false, GGML_LOG_LEVEL_ERROR, nullptr });
    vision_ctx * gpu = vision_init(fname, { /*use_gpu*/ true,  GGML_LOG_LEVEL_ERROR, nullptr });
    if (!cpu || !gpu) { printf("{\"error\": \"vision_init failed\"}\n"); return 1; }
    const int n_tok = vision_n_output_tokens(cpu), n_embd = vision_n_mmproj_embd(cpu);
    Lcg rng{ 777 };
    Cmp cmp;
    std::vector<double> ms_cpu, ms_gpu;
    for (int f = 0; f < n_frames; ++f) {
        vision_image_f32 img;
        img.nx = w; img.ny = h; img.buf.resize((size_t) 3 * w * h);
        for (auto & v : img.buf) v = rng.next();                                           // (pixel / 255 - 0.5) / 0.5 is in [-1, 1]
        std::vector<float> out_cpu((size_t) n_tok * n_embd), out_gpu(out_cpu.size());
        double t0 = now_ms();
        if (!vision_image_encode(cpu, n_threads, &img, out_cpu.data())) { printf("{\"error\": \"cpu encode failed at frame %d\"}\n", f); return 1; }
        double t1 = now_ms();
        if (!vision_image_encode(gpu, n_threads, &img, out_gpu.data())) { printf("{\"error\": \"gpu encode failed at frame %d\"}\n", f); return 1; }
        double t2 = now_ms();
        ms_cpu.push_back(t1 - t0); ms_gpu.push_back(t2 - t1);
        compare(cmp, out_cpu, out_gpu);
    }
    printf("{\"encoder\": \"vpm\", \"frames\": %d, \"image\": [%d, %d], \"tokens_per_frame\": %d, \"n_embd\": %d, \"threads\": %d, \"max_rel_err\": %.3e, \"nmse\": %.3e, "
           "\"non_finite\": %zu, \"ms_cpu\": %s, \"ms_gpu\": %s}\n", n_frames, w, h, n_tok, n_embd, n_threads, cmp.max_diff / fmax(cmp.max_ref, 1e-30),
           cmp.sse / fmax(cmp.ssr, 1e-30), cmp.n_bad, join(ms_cpu).c_str(), join(ms_gpu).c_str());
    vision_free(gpu); vision_free(cpu);
    return 0;
}

int main(int argc, char ** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s apm|vpm model.gguf n [n_threads] [w h]\n", argv[0]); return 2; }
    if (!getenv("OMNI_NO_LOAD_ALL")) ggml_backend_load_all();
    const int n = atoi(argv[3]), n_threads = argc > 4 ? atoi(argv[4]) : 8;
    std::string devs;
    for (size_t i = 0; i < ggml_backend_dev_count(); ++i) devs += std::string(i ? ", " : "") + ggml_backend_dev_name(ggml_backend_dev_get(i));
    fprintf(stderr, "devices: %s\n", devs.c_str());
    if (strcmp(argv[1], "apm") == 0) return run_apm(argv[2], n, n_threads);
    if (strcmp(argv[1], "vpm") == 0) return run_vpm(argv[2], n, n_threads, argc > 6 ? atoi(argv[5]) : 448, argc > 6 ? atoi(argv[6]) : 448);
    fprintf(stderr, "unknown encoder %s\n", argv[1]);
    return 2;
}
