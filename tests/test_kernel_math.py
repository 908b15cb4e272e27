"""CPU: host-checkable properties of arithmetic / layout tricks the sm_100a kernels rely on (numpy restatements of the device code, so that a
change of the formula in csrc/ has to be made here too)."""
import numpy as np


def nearest_int_magic(x: np.ndarray) -> np.ndarray:
    """csrc/quant_dev.cuh nearest_int_magic == the reference's nearest_int (ggml-quants.c): add 1.5 * 2^23 in f32, read the mantissa."""
    y = (x.astype(np.float32) + np.float32(12582912.0)).astype(np.float32)
    return y.view(np.int32) - np.int32(0x4B400000)


def test_nearest_int_magic_is_round_half_even():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-127.6, 127.6, 200000), np.arange(-128, 128) + 0.5, np.arange(-128, 128) - 0.5,
                        rng.uniform(-4e6, 4e6, 50000)]).astype(np.float32)
    assert np.array_equal(nearest_int_magic(x), np.rint(x.astype(np.float64)).astype(np.int32))


def act_qs_off(off: np.ndarray) -> np.ndarray:
    """csrc/quant_dev.cuh act_qs_off<true>: 16-byte segment i of the r-th 128-byte region lives at position (i + r) & 7."""
    seg = off >> 4
    r = seg >> 3
    return (off & 15) | ((r * 8 + ((seg + r) & 7)) << 4)


def test_activation_record_rotation_is_a_bijection_and_conflict_free():
    k = 12288
    off = np.arange(k, dtype=np.int64)
    phys = act_qs_off(off)
    assert np.array_equal(np.sort(phys), off)                              # a permutation of the record's bytes
    assert np.array_equal(phys >> 7, off >> 7)                             # that never leaves its 128-byte region
    # hfrag_fill (csrc/stream_dot.cuh): lane l reads the 8 segments of region l, one per LDS.128; a quarter-warp (8 consecutive lanes) must hit
    # 8 different 16-byte bank groups (bank group = bits 4..6 of the shared-memory address)
    for i in range(8):
        for q in range(4):
            lanes = np.arange(8 * q, 8 * q + 8, dtype=np.int64)
            groups = (act_qs_off(lanes * 128 + 16 * i) >> 4) & 7
            assert len(set(groups.tolist())) == 8, (i, q, groups)
    # the quantiser's 8-byte stores (lane owns bytes [8 lane, 8 lane + 8) of a 256-byte block) stay inside one 16-byte segment
    st = act_qs_off(np.arange(0, 4096, 8, dtype=np.int64))
    assert np.all(st % 8 == 0) and np.all((st >> 4) == (act_qs_off(np.arange(7, 4096, 8, dtype=np.int64)) >> 4))


def umma_desc(saddr: int, lbo: int, sbo: int, layout: int = 0) -> int:
    """csrc/mmq_tc.cu tc_desc / tc_desc_sw128: cute::UMMA::SmemDescriptor bit fields (start >> 4 at 0, LBO >> 4 at 16, SBO >> 4 at 32,
    version 1 at 46, layout type at 61)."""
    return ((saddr & 0x3ffff) >> 4) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (layout << 61)


def test_umma_descriptor_fields_fit_and_round_trip():
    for saddr, lbo, sbo, layout in ((0, 2048, 128, 0), (147456 + 16384, 4096, 128, 0), (226 * 1024, 16, 1024, 2)):
        d = umma_desc(saddr, lbo, sbo, layout)
        assert d < 1 << 64
        assert (d & 0x3fff) << 4 == saddr and ((d >> 16) & 0x3fff) << 4 == lbo and ((d >> 32) & 0x3fff) << 4 == sbo
        assert (d >> 46) & 3 == 1 and d >> 61 == layout
    # instruction descriptor (tc_idesc): D = F32 at bit 4, N >> 3 at bit 17 (6 bits), M >> 4 at bit 24 (5 bits)
    for n in range(16, 257, 16):
        idesc = (1 << 4) | ((n >> 3) << 17) | ((128 >> 4) << 24)
        assert (idesc >> 17) & 0x3f == n >> 3 and (idesc >> 24) & 0x1f == 8 and idesc < 1 << 32
