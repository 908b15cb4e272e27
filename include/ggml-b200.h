/* include/ggml-b200.h — entry points of the ggml backend plugin libggml-b200.so (the drop-in boundary).
 *
 * The reference (llama.cpp-omni / ggml 0.9.4) discovers backends through its registry:
 *   - dynamic:  GGML_BACKEND_PATH=/path/libggml-b200.so  ->  ggml_backend_load_all() dlopens the library and looks up
 *               `ggml_backend_init` (required) and `ggml_backend_score` (optional) with dlsym
 *               (ggml/src/ggml-backend-reg.cpp:257-273, 581-608); reg->api_version must be GGML_BACKEND_API_VERSION (2,
 *               ggml/src/ggml-backend-impl.h:11, 206-210);
 *   - static:   a host program that links the library can call ggml_backend_register(ggml_backend_b200_reg())
 *               (ggml/include/ggml-backend.h:218) — what an LD_PRELOAD constructor does for llama-omni-cli, which never
 *               calls ggml_backend_load_all() before loading the LLM.
 * These replace ggml_backend_cuda_reg / ggml_backend_cuda_init of the reference CUDA backend (ggml/include/ggml-cuda.h:22-44).
 *
 * The plugin itself is host C++ (csrc/ggml_b200/ggml-b200.cpp) over the C-ABI kernel library, include/b200_ops.h.
 */
#ifndef GGML_B200_H
#define GGML_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GGML_B200_API __attribute__((visibility("default")))

struct ggml_backend_reg;
struct ggml_backend;

/* dlsym entry points (ggml-backend-reg.cpp:257-273) */
GGML_B200_API struct ggml_backend_reg * ggml_backend_init(void);
GGML_B200_API int                       ggml_backend_score(void);      /* 100 when an sm_100 device is present, else 0 */

/* direct-link equivalents of ggml_backend_cuda_reg() / ggml_backend_cuda_init(device) (ggml-cuda.h:22-30) */
GGML_B200_API struct ggml_backend_reg * ggml_backend_b200_reg(void);
GGML_B200_API struct ggml_backend *     ggml_backend_b200_init(int device);

/* ABI version of the kernel library this plugin was built against (b200_abi_version) */
GGML_B200_API int ggml_backend_b200_abi(void);

#ifdef __cplusplus
}
#endif
#endif /* GGML_B200_H */
