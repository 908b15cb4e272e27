// tests/native/projector_graph.cpp — TEST INFRASTRUCTURE.  The omni projector path (tools/omni/omni.cpp:1068-1260): a graph
//   out = linear2 . relu(linear1 . x + b1) + b2
// built with the ggml API, weights allocated with the backend's DEFAULT buffer type, and run with a DIRECT ggml_backend_graph_compute on
// ggml_backend_init_by_type(GPU) — no scheduler, so no CPU fallback: every node must be supported by the backend or the call fails.
// The same graph runs on the reference CPU backend; prints one JSON line with the NMSE between the two.
//   projector_graph <in_dim> <hid_dim> <out_dim> <n_tokens> <weight type: f32|f16>
#include "ggml.h"
#include "ggml-alloc.h"
#include "ggml-backend.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static std::vector<float> run(ggml_backend_t backend, int in_dim, int hid, int out_dim, int n_tok, ggml_type wt, const std::vector<float> & w1, const std::vector<float> & b1,
                              const std::vector<float> & w2, const std::vector<float> & b2, const std::vector<float> & x, std::string & err) {
    ggml_init_params ip = { ggml_tensor_overhead() * 8, nullptr, true };
    ggml_context * wctx = ggml_init(ip);
    ggml_tensor * W1 = ggml_new_tensor_2d(wctx, wt, in_dim, hid), * B1 = ggml_new_tensor_1d(wctx, GGML_TYPE_F32, hid);
    ggml_tensor * W2 = ggml_new_tensor_2d(wctx, wt, hid, out_dim), * B2 = ggml_new_tensor_1d(wctx, GGML_TYPE_F32, out_dim);
    ggml_backend_buffer_t wbuf = ggml_backend_alloc_ctx_tensors_from_buft(wctx, ggml_backend_get_default_buffer_type(backend));
    if (!wbuf) { err = "weight buffer"; return {}; }
    auto put = [&](ggml_tensor * t, const std::vector<float> & v) {
        if (t->type == GGML_TYPE_F32) { ggml_backend_tensor_set(t, v.data(), 0, v.size() * 4); return; }
        std::vector<ggml_fp16_t> h(v.size());
        ggml_fp32_to_fp16_row(v.data(), h.data(), (int64_t) v.size());
        ggml_backend_tensor_set(t, h.data(), 0, h.size() * 2);
    };
    put(W1, w1); put(B1, b1); put(W2, w2); put(B2, b2);
    ggml_init_params gp = { ggml_tensor_overhead() * 10 + ggml_graph_overhead(), nullptr, true };
    ggml_context * ctx = ggml_init(gp);
    ggml_tensor * input = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, in_dim, n_tok);
    ggml_set_input(input);
    ggml_cgraph * gf = ggml_new_graph(ctx);
    ggml_tensor * hidden = ggml_relu(ctx, ggml_add(ctx, ggml_mul_mat(ctx, W1, input), B1));
    ggml_tensor * output = ggml_add(ctx, ggml_mul_mat(ctx, W2, hidden), B2);
    ggml_build_forward_expand(gf, output);
    ggml_backend_buffer_t cbuf = ggml_backend_alloc_ctx_tensors(ctx, backend);
    if (!cbuf) { err = "compute buffer"; return {}; }
    ggml_backend_tensor_set(input, x.data(), 0, x.size() * 4);
    if (ggml_backend_graph_compute(backend, gf) != GGML_STATUS_SUCCESS) { err = "graph_compute failed (an op of the projector graph is not supported)"; return {}; }
    std::vector<float> out((size_t) out_dim * n_tok);
    ggml_backend_tensor_get(output, out.data(), 0, out.size() * 4);
    ggml_backend_buffer_free(cbuf); ggml_backend_buffer_free(wbuf); ggml_free(ctx); ggml_free(wctx);
    return out;
}

int main(int argc, char ** argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s in hid out n_tokens f32|f16\n", argv[0]); return 2; }
    const int in_dim = atoi(argv[1]), hid = atoi(argv[2]), out_dim = atoi(argv[3]), n_tok = atoi(argv[4]);
    const ggml_type wt = strcmp(argv[5], "f16") == 0 ? GGML_TYPE_F16 : GGML_TYPE_F32;
    ggml_backend_load_all();
    uint32_t s = 777;
    auto rnd = [&](size_t n, float scale) { std::vector<float> v(n); for (auto & f : v) { s = s * 1664525u + 1013904223u; f = scale * ((float) (s >> 8) / 8388608.0f - 1.0f); } return v; };
    const auto w1 = rnd((size_t) in_dim * hid, 0.05f), b1 = rnd(hid, 0.1f), w2 = rnd((size_t) hid * out_dim, 0.05f), b2 = rnd(out_dim, 0.1f), x = rnd((size_t) in_dim * n_tok, 1.0f);
    ggml_backend_t gpu = ggml_backend_init_by_type(GGML_BACKEND_DEVICE_TYPE_GPU, nullptr), cpu = ggml_backend_init_by_type(GGML_BACKEND_DEVICE_TYPE_CPU, nullptr);
    if (!gpu || !cpu) { printf("{\"error\": \"no %s backend\"}\n", gpu ? "CPU" : "GPU"); return 1; }
    std::string err;
    const auto a = run(cpu, in_dim, hid, out_dim, n_tok, wt, w1, b1, w2, b2, x, err);
    const auto b = err.empty() ? run(gpu, in_dim, hid, out_dim, n_tok, wt, w1, b1, w2, b2, x, err) : std::vector<float>();
    if (!err.empty()) { printf("{\"error\": \"%s\"}\n", err.c_str()); return 1; }
    double num = 0, den = 0; bool finite = true;
    for (size_t i = 0; i < a.size(); ++i) { num += ((double) a[i] - b[i]) * ((double) a[i] - b[i]); den += (double) a[i] * a[i]; finite = finite && std::isfinite(b[i]); }
    printf("{\"gpu_backend\": \"%s\", \"nmse\": %.3g, \"finite\": %s, \"n\": %zu}\n", ggml_backend_name(gpu), den > 0 ? num / den : 0.0, finite ? "true" : "false", a.size());
    ggml_backend_free(gpu); ggml_backend_free(cpu);
    return finite && den > 0 && num / den <= 1e-6 ? 0 : 1;
}
