// libggml-b200-preload.so — registers the B200 backend from a load-time constructor, for host programs that never call ggml_backend_load_all():
//   LD_PRELOAD=/path/libggml-b200-preload.so llama-omni-cli -m ... --omni ...
// tools/omni/omni-cli.cpp:198-350 parses its own arguments and loads the LLM without touching the dynamic backend loader (only token2wav-impl.cpp:6287 calls it,
// later), so GGML_BACKEND_PATH alone never reaches it.  ggml_backend_register (ggml/include/ggml-backend.h:218) appends to the registry, a
// function-local static (ggml/src/ggml-backend-reg.cpp:320-323), so calling it before main() is safe.  Do not ALSO set GGML_BACKEND_PATH for tools that do call
// ggml_backend_load_all(): the device would be listed twice.
#include "ggml-backend.h"
#include "../../../include/ggml-b200.h"

__attribute__((constructor)) static void ggml_b200_preload(void) {
    if (ggml_backend_score() > 0) ggml_backend_register((ggml_backend_reg_t) ggml_backend_b200_reg());
}
