"""llama.cpp-omni_b200 — B200-native ggml backend for the llama.cpp-omni LLM hot path.

The product is two native libraries built from csrc/ (see DESIGN.md):
  lib/libb200ops.so    hand-written sm_100a kernels behind the C-ABI of include/b200_ops.h
  lib/libggml-b200.so  the ggml backend plugin (ggml_backend_init) the reference loads with GGML_BACKEND_PATH
This Python package is only the thin ctypes mirror of that C-ABI used by tests/, bench.py and __graft_entry__.py; torch
supplies device memory and streams.  There is NO CPU fallback: importing `ops` without the built library raises.

The directory name is not a valid Python identifier; load it with `__graft_entry__.load_package()`.
"""
from . import ops, decode, pipeline  # noqa: F401
