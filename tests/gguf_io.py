"""Minimal GGUF v3 reader (TEST INFRASTRUCTURE): just enough of the format read by ggml/src/gguf.cpp to pull the raw tensor bytes out of a
file the reference's own llama-quantize wrote, so that the CPU oracle chain (tests/oracle_decode.py) can run on realistically quantised weights."""
from __future__ import annotations

import struct

import numpy as np

BLOCK = {0: (1, 4), 1: (1, 2), 2: (32, 18), 8: (32, 34), 12: (256, 144), 13: (256, 176), 14: (256, 210)}
_SCALAR = {0: "B", 1: "b", 2: "H", 3: "h", 4: "I", 5: "i", 6: "f", 7: "?", 10: "Q", 11: "q", 12: "d"}


def read_gguf(path: str) -> tuple[dict, dict]:
    """-> (metadata {key: value}, tensors {name: (ggml_type, [ne0, ne1, ...], uint8 array of the raw blocks)})."""
    buf = np.memmap(path, dtype=np.uint8, mode="r")
    off = 0

    def rd(fmt):
        nonlocal off
        v = struct.unpack_from("<" + fmt, buf, off)
        off += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v

    def rstr():
        nonlocal off
        n = rd("Q")
        s = bytes(buf[off:off + n]).decode("utf-8", "replace")
        off += n
        return s

    def rval(t):
        if t in _SCALAR:
            return rd(_SCALAR[t])
        if t == 8:
            return rstr()
        if t == 9:
            et, n = rd("I"), rd("Q")
            return [rval(et) for _ in range(n)]
        raise ValueError(f"gguf value type {t}")

    magic, version, n_tensors, n_kv = rd("I"), rd("I"), rd("Q"), rd("Q")
    assert magic == 0x46554747 and version == 3, (hex(magic), version)
    meta = {}
    for _ in range(n_kv):
        k = rstr()
        meta[k] = rval(rd("I"))
    infos = []
    for _ in range(n_tensors):
        name = rstr()
        nd = rd("I")
        ne = [rd("Q") for _ in range(nd)]
        t, o = rd("I"), rd("Q")
        infos.append((name, t, ne, o))
    align = int(meta.get("general.alignment", 32))
    data0 = (off + align - 1) // align * align
    tensors = {}
    for name, t, ne, o in infos:
        qk, bs = BLOCK[t]
        n = int(np.prod(ne))
        nbytes = n // qk * bs
        tensors[name] = (t, ne, np.array(buf[data0 + o:data0 + o + nbytes]))
    return meta, tensors
