"""BASELINE.json configs[3], the encoder half ("APM 1 s audio chunk + VPM frame"): the reference's UNMODIFIED tools/omni/audition.cpp and
vision.cpp, compiled from where they lie by oracle/Makefile into oracle/_ref/bin/omni_encoders (tests/native/omni_encoders.cpp), on synthetic GGUFs written by
tools/make_omni_gguf.py with the KV keys / tensor names their loaders ask for.

  * CPU (not gpu): the GGUFs load in the reference's own loaders and both encoders run on the reference CPU backend (the harness's two sides are then the same backend:
    outputs identical and finite) — the test of the GGUF writer and of the harness;
  * GPU: one side is the reference CPU backend, the other whatever GPU-type backend is registered = libggml-b200.so, through each encoder's own
    ggml_backend_sched(GPU, CPU); embeddings must agree within the F16-GEMM tolerance, through GGML_BACKEND_PATH and through the LD_PRELOAD shim (no ggml_backend_load_all()
    in the host program, as llama-omni-cli).
(The file name sorts last on purpose: `pytest -x` reaches these after the kernel and plugin tests.)"""
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
EXE = REF / "bin" / "omni_encoders"
LIB = ROOT / "llama.cpp-omni_b200" / "lib"
pytestmark = pytest.mark.skipif(not EXE.exists(), reason="oracle/_ref/bin/omni_encoders is not built (needs /root/reference)")

SMALL = {"apm": ["--layers", "2", "--d-model", "256", "--heads", "4", "--proj", "512"],
         "vpm": ["--layers", "2", "--embd", "288", "--heads", "4", "--ff", "512", "--proj", "512"]}


_CACHE = {}


@pytest.fixture(scope="module", autouse=True)
def _remove_the_ggufs_afterwards():
    yield
    import shutil
    for f in _CACHE.values():
        shutil.rmtree(f.parent, ignore_errors=True)
    _CACHE.clear()


def _gguf(tmp_path, what, full=False):
    """one file per (encoder, size) and session: the full-size ones are 0.7 / 1.1 GB"""
    if (what, full) not in _CACHE:
        f = Path(tempfile.mkdtemp(prefix="omni_gguf_")) / f"{what}.gguf"
        subprocess.check_call([sys.executable, str(ROOT / "tools" / "make_omni_gguf.py"), what, str(f)] + ([] if full else SMALL[what]), stderr=subprocess.DEVNULL)
        _CACHE[(what, full)] = f
    return _CACHE[(what, full)]


def _run(args, **extra):
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = f"{REF / 'lib'}:{LIB}:" + env.get("LD_LIBRARY_PATH", "")
    env.pop("GGML_BACKEND_PATH", None)
    env.update(extra)
    r = subprocess.run([str(EXE)] + [str(a) for a in args], env=env, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, (r.returncode, r.stdout[-500:], r.stderr[-1500:])
    return json.loads(lines[-1]), r.stderr


@pytest.mark.parametrize("what,n", [("apm", 3), ("vpm", 1)])
def test_synthetic_encoder_ggufs_load_and_run_in_the_reference(tmp_path, what, n):
    res, err = _run([what, _gguf(tmp_path, what), n, 4])
    assert "error" not in res, (res, err[-800:])
    assert res["non_finite"] == 0 and res["max_rel_err"] == 0.0                     # no GPU-type device here: both sides are the reference CPU backend
    assert res["tokens_per_chunk" if what == "apm" else "tokens_per_frame"] == (10 if what == "apm" else 64) and res["n_embd"] == 512
    assert "devices: CPU" in err


def test_sched_census_parses_the_reference_scheduler_dump(tmp_path):
    sys.path.insert(0, str(ROOT / "tools"))
    import sched_census
    _, err = _run(["apm", _gguf(tmp_path, "apm"), 1, 2], GGML_SCHED_DEBUG="2")
    c = sched_census.census(err)
    assert c["splits"] == {"CPU": 1} and c["nodes"]["CPU"]["MUL_MAT"] == 20 and c["nodes"]["CPU"]["POOL_1D"] == 1, c


# the bars: F16 weights with F16-rounded activations and F32 accumulation on both sides (ggml-cpu vec_dot_f16 vs tcgen05 kind::f16) — what differs is the summation order
# and the F16 rounding points of intermediate activations; measured values are recorded in profiles/r02_omni_encoders.md
@pytest.mark.gpu
@pytest.mark.parametrize("what,n,route", [("apm", 4, "path"), ("vpm", 1, "path"), ("apm", 2, "preload")])
def test_omni_encoders_on_the_plugin_match_the_reference_cpu_backend(tmp_path, what, n, route):
    env = {"GGML_BACKEND_PATH": str(LIB / "libggml-b200.so")} if route == "path" else {"OMNI_NO_LOAD_ALL": "1", "LD_PRELOAD": str(LIB / "libggml-b200-preload.so")}
    res, err = _run([what, _gguf(tmp_path, what, full=True), n, min(os.cpu_count() or 8, 16)], **env)
    assert "error" not in res, (res, err[-800:])
    assert "B200" in err.split("devices:")[-1].splitlines()[0], err[-500:]
    assert res["non_finite"] == 0 and res["nmse"] <= 1e-6 and res["max_rel_err"] <= 2e-3, res          # measured: 3e-4 / 8e-8 (profiles/r02_omni_encoders.md)
    out = os.environ.get("OMNI_RESULTS_DIR")
    if out:
        (Path(out) / f"r02_omni_{what}_{route}.json").write_text(json.dumps(res) + "\n")


@pytest.mark.gpu
@pytest.mark.parametrize("what", ["apm", "vpm"])
def test_every_node_of_the_encoder_graphs_runs_on_the_plugin(tmp_path, what):
    """GGML_SCHED_DEBUG=2 makes the reference scheduler print where it placed every node (ggml-backend.cpp:843-881).  Nothing may be left on the CPU backend: every CPU
    split inside a device graph costs a device synchronisation and two copies.  (Before k_mm_simt took F16 activations, the F16 x F16 products of ggml_conv_1d /
    conv_2d — im2col matrix x kernel — were the only nodes left there: 2 per audio chunk, 1 per frame.)"""
    sys.path.insert(0, str(ROOT / "tools"))
    import sched_census
    res, err = _run([what, _gguf(tmp_path, what, full=True), 1, min(os.cpu_count() or 8, 16)], GGML_BACKEND_PATH=str(LIB / "libggml-b200.so"), GGML_SCHED_DEBUG="2")
    assert "error" not in res, (res, err[-800:])
    c = sched_census.census(err)
    assert "error" not in c, err[-1500:]
    assert not c["nodes"].get("CPU"), c
    assert list(c["splits"]) and all(b.startswith("B200") for b in c["splits"]), c
    out = os.environ.get("OMNI_RESULTS_DIR")
    if out:
        (Path(out) / f"r02_omni_{what}_sched_census.json").write_text(json.dumps(c) + "\n")
