"""The lane state machines of the coder kernels (icer_compression_b200/csrc/coder_core.h), built for the HOST (tests/emul) and compared
with the oracle's segment-plane coder (oracle/icer_oracle.c restating icer_context_modeller.c:312-602, icer_encoding.c,
icer_decoding.c).  CPU only: this is the part of the CUDA path that can be checked without a GPU -- the bit-slice representation,
the elastic per-lane stepping, quiet-run skipping, the interleave window incl. forced flushes, and the decoder's halo rules."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import harness

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL_DIR = os.path.join(HERE, "emul")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libemul_coder.so")
    src = os.path.join(EMUL_DIR, "emul_coder.cpp")
    core = os.path.join(HERE, "..", "icer_compression_b200", "csrc")
    newest = max(os.path.getmtime(p) for p in [src, os.path.join(core, "coder_core.h"), os.path.join(core, "coder_tables.h")])
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    lib = C.CDLL(so)
    lib.emul_encode_segment_plane.argtypes = [C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                              C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_uint32)]
    lib.emul_encode_segment_plane.restype = C.c_long
    lib.emul_decode_segment_plane.argtypes = [C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                              C.POINTER(C.c_uint8), C.c_size_t, C.c_uint32]
    return lib


@pytest.fixture(scope="module")
def orc():
    o = harness.Oracle()
    o.lib.orc_decode_segment_plane.argtypes = [C.POINTER(C.c_uint16), C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                               C.POINTER(C.c_uint8), C.c_uint32]
    return o


def _p16(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint16))


def _p8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def make_segment(rng, w, h, kind):
    if kind == "sparse":
        mag = (rng.random((h, w)) < 0.05) * rng.integers(1, 600, (h, w))
    elif kind == "laplace":
        mag = np.minimum(np.abs(rng.laplace(0, 6, (h, w))).astype(np.int64), 0x7FFF)
    elif kind == "noise":
        mag = rng.integers(0, 512, (h, w))
    elif kind == "wide":
        mag = rng.integers(0, 0x8000, (h, w))
    else:
        mag = np.zeros((h, w), dtype=np.int64)
    sign = rng.integers(0, 2, (h, w)) * (mag != 0)
    return (mag | (sign << 15)).astype(np.uint16)


def check_plane(emul, orc, seg, sb, lsb):
    h, w = seg.shape
    cap = w * h * 3 + 64
    a = np.ascontiguousarray(seg)
    out_o = np.zeros(cap + 8, dtype=np.uint8)
    out_e = np.zeros(cap + 8, dtype=np.uint8)
    nb_o = orc.lib.orc_encode_segment_plane(_p16(a), w, h, w, sb, lsb, _p8(out_o), cap)
    steps = C.c_uint32(0)
    nb_e = emul.emul_encode_segment_plane(_p16(a), w, h, w, sb, lsb, _p8(out_e), cap, C.byref(steps))
    assert nb_o >= 0
    assert nb_e == nb_o, "bit count %d vs oracle %d (w=%d h=%d sb=%d lsb=%d)" % (nb_e, nb_o, w, h, sb, lsb)
    nbytes = (nb_o + 7) // 8
    assert bytes(out_e[:nbytes]) == bytes(out_o[:nbytes]), "payload differs (w=%d h=%d sb=%d lsb=%d)" % (w, h, sb, lsb)
    # decode: the planes above lsb are known, the payload is the oracle's
    above = (a & np.uint16(0x7FFF)) >> np.uint16(lsb + 1) << np.uint16(lsb + 1)
    above = (above | np.where(above != 0, a & np.uint16(0x8000), 0)).astype(np.uint16)
    d_o = above.copy()
    d_e = above.copy()
    rc_o = orc.lib.orc_decode_segment_plane(_p16(d_o), w, h, w, sb, lsb, _p8(out_o), nb_o)
    rc_e = emul.emul_decode_segment_plane(_p16(d_e), w, h, w, sb, lsb, _p8(out_e), nbytes, nb_o)
    assert rc_e == rc_o
    assert (d_e == d_o).all(), "decoded plane differs (w=%d h=%d sb=%d lsb=%d)" % (w, h, sb, lsb)
    return steps.value


@pytest.mark.parametrize("kind", ["zero", "sparse", "laplace", "noise", "wide"])
def test_lane_machines_vs_oracle_small_segments(emul, orc, kind):
    rng = np.random.default_rng(1234)
    shapes = [(1, 1), (2, 1), (1, 5), (3, 3), (29, 2), (30, 3), (31, 3), (59, 4), (60, 2), (61, 5), (90, 3), (97, 7), (128, 9), (7, 40)]
    for (w, h) in shapes:
        seg = make_segment(rng, w, h, kind)
        for sb in range(4):
            for lsb in ([0, 3, 8] if kind != "wide" else [0, 5, 11, 14]):
                check_plane(emul, orc, seg, sb, lsb)


def test_lane_machines_forced_flush_and_long_runs(emul, orc):
    """A large noise segment keeps more than 2048 codewords outstanding (forced flushes, icer_encoding.c:59-64); a large, almost empty
    one makes the quiet fast path cross every kind of event (codeword complete, bin change, count rescale)."""
    rng = np.random.default_rng(7)
    seg = make_segment(rng, 256, 120, "noise")
    for lsb in (0, 4, 8):
        check_plane(emul, orc, seg, 3, lsb)
    seg = make_segment(rng, 300, 150, "sparse")
    seg[40:90] = 0
    for sb in (0, 1):
        for lsb in (0, 2, 9):
            check_plane(emul, orc, seg, sb, lsb)


def test_lane_machines_on_a_transformed_frame(emul, orc):
    """every segment plane of a real coefficient image (the benchmark generator, 6 stages)"""
    img = orc.synth(192, 160, 3, 20260923, 0)
    co = orc.compress([img], 4, 0, 6, 192 * 160)["coefs"][0]
    for (x, y, w, h, sb) in [(96, 0, 96, 80, 1), (0, 80, 96, 80, 2), (96, 80, 96, 80, 3), (0, 0, 12, 10, 0), (48, 40, 48, 40, 3)]:
        seg = np.ascontiguousarray(co[y:y + h, x:x + w])
        for lsb in range(9):
            check_plane(emul, orc, seg, sb, lsb)


def test_lane_machines_random_segments_property(emul, orc):
    """hypothesis: random segment shapes (incl. widths around the 30-pixel chunk size), densities, sub-band kinds and planes"""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=120, deadline=None, suppress_health_check=list(HealthCheck))
    @given(w=st.integers(1, 125), h=st.integers(1, 9), sb=st.integers(0, 3), lsb=st.integers(0, 14), seed=st.integers(0, 2 ** 31 - 1),
           density=st.sampled_from([0.0, 0.02, 0.2, 0.7, 1.0]), maxbits=st.integers(1, 15))
    def run(w, h, sb, lsb, seed, density, maxbits):
        rng = np.random.default_rng(seed)
        mag = (rng.random((h, w)) < density) * rng.integers(0, 1 << maxbits, (h, w))
        sign = rng.integers(0, 2, (h, w)) * (mag != 0)
        seg = (mag | (sign << 15)).astype(np.uint16)
        check_plane(emul, orc, seg, sb, lsb)

    run()
