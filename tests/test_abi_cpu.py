"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that include/*.h
declares (no compute calls: there is no GPU here), and the host-only helpers match the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import harness

ROOT = harness.ROOT


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"static inline[^{]*\{[^}]*\}", "", txt)
    return sorted(set(re.findall(r"\b(icer_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    import icer_compression_b200 as icer
    return icer.load_library()


@pytest.mark.parametrize("header", ["icer.h", "icer_b200.h"])
def test_every_declared_symbol_is_exported(lib, header):
    names = declared_functions(header)
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "%s declared in include/%s but not exported" % (n, header)


def test_no_oracle_linkage():
    """The product must not link or reference anything under oracle/."""
    import subprocess
    import icer_compression_b200 as icer
    out = subprocess.check_output(["ldd", icer.lib_path()]).decode()
    assert "oracle" not in out and "icer_ref" not in out
    syms = subprocess.check_output(["nm", "-D", icer.lib_path()]).decode()
    assert "orc_" not in syms


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import icer_compression_b200 as icer
    with pytest.raises(icer.IcerError):
        icer.IcerB200()
    assert icer.load_library().icer_init() == icer.ICER_FATAL_ERROR


def test_partition_parameters_match_reference(lib, ref9):
    class PP(C.Structure):
        _fields_ = [(n, C.c_uint16) for n in "w h r c r_t h_t x_t c_t0 y_t r_t0 x_b c_b0 y_b r_b0 s".split()]
    lib.icer_generate_partition_parameters.argtypes = [C.POINTER(PP), C.c_size_t, C.c_size_t, C.c_uint16]
    ref9.lib.icer_generate_partition_parameters.argtypes = [C.POINTER(PP), C.c_size_t, C.c_size_t, C.c_uint16]
    rng = np.random.default_rng(0)
    for _ in range(400):
        w, h, s = int(rng.integers(1, 700)), int(rng.integers(1, 700)), int(rng.integers(1, 40))
        a, b = PP(), PP()
        ra = lib.icer_generate_partition_parameters(C.byref(a), w, h, s)
        rb = ref9.lib.icer_generate_partition_parameters(C.byref(b), w, h, s)
        assert ra == rb, (w, h, s)
        if ra == 0:
            assert bytes(a) == bytes(b), (w, h, s)


def test_host_scanner_matches_reference(lib, ref9):
    """icer_find_packet_in_bytestream / icer_get_image_dimensions (host code in the product) vs the reference."""
    img = harness.load_boat()[:96, :128].copy()
    s = ref9.compress([img], 3, 0, 4, 96 * 128)["stream"]
    rng = np.random.default_rng(5)
    variants = [s, s[10:], s[:len(s) // 3]]
    for _ in range(4):
        t = bytearray(s)
        for pos in rng.integers(0, len(t), 12):
            t[pos] ^= 0x5A
        variants.append(bytes(t))
    u8p, szp = C.POINTER(C.c_uint8), C.POINTER(C.c_size_t)
    for L in (lib, ref9.lib):
        L.icer_find_packet_in_bytestream.argtypes = [C.POINTER(C.c_void_p), u8p, C.c_size_t, szp]
        L.icer_get_image_dimensions.argtypes = [u8p, C.c_size_t, szp, szp]
    for v in variants:
        buf = np.frombuffer(v + b"\0" * 64, dtype=np.uint8).copy()
        p = buf.ctypes.data_as(u8p)
        res = []
        for L in (lib, ref9.lib):
            offs, off = [], 0
            while off < len(v):
                seg, adv = C.c_void_p(), C.c_size_t(0)
                q = C.cast(C.c_void_p(buf.ctypes.data + off), u8p)
                rc = L.icer_find_packet_in_bytestream(C.byref(seg), q, len(v) - off, C.byref(adv))
                offs.append((rc, adv.value, (seg.value - buf.ctypes.data) if seg.value else None))
                off += adv.value
            w, h = C.c_size_t(0), C.c_size_t(0)
            rc = L.icer_get_image_dimensions(p, len(v), C.byref(w), C.byref(h))
            res.append((offs, rc, (w.value, h.value) if rc == 0 else None))
        assert res[0] == res[1]


def test_u8_interleave_map_matches_reference(ref9):
    """The native uint8 decoder has to reproduce where lib_icer's icer_interleave_uint8 really puts every element (not the
    perfect shuffle for odd lengths).  The library derives that placement on the host; check it against the
    compiled reference for every length the transform can meet in small images and a sample of large ones."""
    import numpy as np
    from icer_compression_b200 import u8_interleave_map
    for n in list(range(2, 1200)) + [2047, 2048, 2049, 4095, 4097, 8191]:
        a = ref9.interleave8_map(n)
        b = u8_interleave_map(n).astype(np.int64)
        assert (a == b).all(), n
        if n % 2 == 0:
            half = n // 2
            assert (b[:half] == 2 * np.arange(half)).all() and (b[half:] == 2 * np.arange(half) + 1).all()
