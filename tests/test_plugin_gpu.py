"""GPU: the drop-in boundary itself.  The UNMODIFIED reference binaries built by oracle/Makefile (test-backend-ops, llama_parity over libllama)
load libggml-b200.so through GGML_BACKEND_PATH, exactly as a llama.cpp-omni user would (ggml-backend-reg.cpp:581-608):
  * the reference's own per-op parity harness (tests/test-backend-ops.cpp: new backend vs the reference CPU backend, its NMSE bars) on the
    hot-path ops;
  * an end-to-end run through libllama on a small Q4_K_M model: the plugin's whole-token decode engine route against its per-op route, and
    against the reference CPU backend, with the reference CPU backend against ITSELF (plain vs repacked weights) as the yardstick for what
    two correct implementations differ by on this model (int8 activation quantisation is discontinuous).
Skipped when oracle/_ref was not built (it is built where /root/reference exists and travels prebuilt to the GPU box)."""
import json
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
PLUGIN = ROOT / "llama.cpp-omni_b200" / "lib" / "libggml-b200.so"
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (REF / "bin" / "test-backend-ops").exists() or not PLUGIN.exists(),
                                                  reason="oracle/_ref or the plugin is not built")]


def _env(plugin=True, **extra):
    e = dict(os.environ)
    e["LD_LIBRARY_PATH"] = f"{REF / 'lib'}:{PLUGIN.parent}:" + e.get("LD_LIBRARY_PATH", "")
    if plugin:
        e["GGML_BACKEND_PATH"] = str(PLUGIN)
    else:
        e.pop("GGML_BACKEND_PATH", None)
    e.update(extra)
    return e


def test_plugin_exports_the_dlsym_entry_points():
    out = subprocess.check_output(["nm", "-D", "--defined-only", str(PLUGIN)], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert {"ggml_backend_init", "ggml_backend_score", "ggml_backend_b200_reg", "ggml_backend_b200_init", "ggml_backend_b200_abi"} <= exported


@pytest.mark.parametrize("op", ["MUL_MAT", "FLASH_ATTN_EXT", "RMS_NORM", "ROPE", "SET_ROWS", "GET_ROWS", "GLU", "ADD", "MUL", "CPY", "SOFT_MAX"])
def test_reference_backend_ops_harness(op):
    r = subprocess.run([str(REF / "bin" / "test-backend-ops"), "test", "-b", "B200:0", "-o", op], env=_env(), capture_output=True, text=True, timeout=900)
    out = re.sub(r"\x1b\[[0-9;]*m", "", r.stdout + r.stderr)
    m = re.search(r"(\d+)/(\d+) tests passed", out)
    assert m, out[-2000:]
    assert m.group(1) == m.group(2) and int(m.group(2)) > 0, out[-3000:]
    assert "Backend B200:0: OK" in out


@pytest.fixture(scope="module")
def small_model(tmp_path_factory):
    d = tmp_path_factory.mktemp("gguf")
    f32, q4 = d / "f32.gguf", d / "q4_k_m.gguf"
    subprocess.check_call([sys.executable, str(ROOT / "tools" / "make_gguf.py"), str(f32), "--layers", "4", "--vocab", "8192", "--ftype", "f32"], timeout=600)
    subprocess.check_call([str(REF / "bin" / "llama-quantize"), str(f32), str(q4), "q4_k_m", str(os.cpu_count() or 4)], env=_env(plugin=False),
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
    f32.unlink()
    return q4


def _parity(model, *args, **env):
    r = subprocess.run([str(REF / "bin" / "llama_parity"), str(model), *map(str, args)], env=_env(**env), capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, (r.stdout + r.stderr)[-2000:]
    return json.loads(lines[-1])


def test_llama_end_to_end_engine_route(small_model):
    threads = os.cpu_count() or 4
    eng = _parity(small_model, 48, 32, threads, 1, 5)                     # mode 5: plugin per-op route (baseline) vs decode-engine route
    assert "error" not in eng, eng
    assert eng["prefill_rel_err"] == 0                                   # same prefill kernels on both sides
    cpu_self = _parity(small_model, 48, 32, threads, 1, 1)               # reference CPU vs reference CPU with repacked weights
    gpu = _parity(small_model, 48, 32, threads, 1)                       # reference CPU vs plugin (engine route)
    yard = max(cpu_self["max_rel_logit_err"], 1e-3)
    # the two routes of the plugin, and the plugin vs the CPU, must differ by no more than ~2x what the reference differs from itself
    assert eng["max_rel_logit_err"] <= 2.0 * yard, (eng, cpu_self)
    assert gpu["max_rel_logit_err"] <= 2.5 * yard, (gpu, cpu_self)
    assert gpu["gpu_tg_tok_s"] > gpu["cpu_tg_tok_s"]
