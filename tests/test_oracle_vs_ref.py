"""Pin the oracle (oracle/liboracle.so) against the UNMODIFIED reference (oracle/_ref/*.so) and against the
committed golden vectors (tests/golden/golden.json, generated from that reference).  CPU only."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

import harness

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))["cases"]


def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()


def case_input(c, orc):
    n = c["name"]
    if n.startswith("boatcolor"):
        return harness.load_boatcolor_yuv()
    if n.startswith("boat"):
        return [harness.load_boat()]
    if n.startswith("synth512_nat8"):
        return [orc.synth(512, 512, 3, 20260922 + 2, 0)]
    if n.startswith("synth512_noise8"):
        return [orc.synth(512, 512, 3, 20260922 + 2, 1)]
    if n.startswith("synth512_yuv"):
        return [orc.synth(512, 512, f, 20260922 + 3, 0) for f in (5, 1005, 2005)]
    if n.startswith("synth320x200"):
        return [orc.synth(320, 200, 1, 20260922 + 4, 2)]
    if n.startswith("synth1024_nat12"):
        return [orc.synth(1024, 1024, 0, 20260922 + 4, 2)]
    if n.startswith("synth301x203"):
        return [orc.synth(301, 203, 7, 99, 0)]
    if n.startswith("synth1024_nat8"):
        return [orc.synth(1024, 1024, 11, 20260922 + 5, 0)]
    raise KeyError(n)


@pytest.mark.parametrize("c", GOLD, ids=[c["name"] for c in GOLD])
def test_oracle_matches_golden(c, oracle9, oracle15):
    orc = oracle15 if c["planes"] == 15 else oracle9
    imgs = case_input(c, orc)
    enc = orc.compress(imgs, c["stages"], c["filt"], c["segments"], c["quota"])
    assert enc["rc"] == c["rc"]
    assert enc["size_used"] == c["size_used"]
    assert md5(enc["stream"]) == c["stream_md5"]
    assert md5(enc["prio"]) == c["prio_md5"]
    assert md5(b"".join(x.tobytes() for x in enc["coefs"])) == c["coefs_md5"]
    dec = orc.decompress(enc["stream"], c["chan"], c["w"] * c["h"], c["stages"], c["filt"], c["segments"])
    assert dec["rc"] == c["dec_rc"]
    assert (dec["w"], dec["h"]) == (c["w"], c["h"])
    assert md5(b"".join(i.tobytes() for i in dec["images"])) == c["decoded_md5"]


def test_tables_match_reference(ref9, oracle9):
    import ctypes as C
    enc = (C.c_uint8 * 3 * 32 * 17)()
    dec = (C.c_uint8 * 3 * 32 * 17)()
    fl = (C.c_uint8 * 2 * 6 * 9 * 17)()
    gol = (C.c_uint16 * 3 * 17)()
    oracle9.lib.orc_get_code_tables(enc, dec, fl, gol)
    r_enc = (C.c_uint8 * 3 * 32 * 17).in_dll(ref9.lib, "icer_custom_coding_scheme")   # {in_bits,out_bits,out}
    r_dec = (C.c_uint8 * 3 * 32 * 17).in_dll(ref9.lib, "icer_custom_decode_scheme")
    r_fl = (C.c_uint8 * 2 * 6 * 9 * 17).in_dll(ref9.lib, "icer_custom_code_flush_bits")
    r_gol = (C.c_uint16 * 3 * 17).in_dll(ref9.lib, "icer_golomb_coders")
    a, b = np.array(enc), np.array(r_enc)
    # the reference leaves out/out_bits of unused entries at 0 as well (BSS)
    assert (a == b).all()
    assert (np.array(dec) == np.array(r_dec)).all()
    assert (np.array(fl) == np.array(r_fl)).all()
    assert (np.array(gol) == np.array(r_gol)).all()


def test_bin_of(ref9, oracle9):
    for t in range(1, 501, 7):
        for z in range((t + 1) // 2, t + 1):
            assert oracle9.bin_of(z, t) == ref9.bin_of(z, t)


def test_crc32(oracle9):
    import zlib
    rng = np.random.default_rng(1)
    for n in (0, 1, 24, 1000):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle9.crc32(d) == zlib.crc32(d)


@pytest.mark.parametrize("filt", range(7))
def test_dwt_1d_all_lengths(filt, ref9, oracle9):
    rng = np.random.default_rng(filt)
    for n in list(range(2, 70)) + [127, 128, 255, 513]:
        for scale in (256, 4096, 65536):
            x = rng.integers(0, scale, n).astype(np.uint16)
            rr, ra = ref9.dwt_1d(x, filt)
            orr, oa = oracle9.dwt_1d(x, filt)
            assert rr == orr and (ra == oa).all(), (n, scale)
            y = rng.integers(0, scale, n).astype(np.uint16)
            rr, ra = ref9.dwt_1d(y, filt, inverse=True)
            orr, oa = oracle9.dwt_1d(y, filt, inverse=True)
            assert rr == orr and (ra == oa).all(), (n, scale, "inv")


@pytest.mark.parametrize("filt", range(7))
def test_dwt_2d_stages(filt, ref9, oracle9):
    rng = np.random.default_rng(100 + filt)
    for (w, h, st) in ((64, 48, 3), (37, 53, 2), (5, 5, 1), (129, 65, 5), (96, 96, 6), (200, 7, 1)):
        img = rng.integers(0, 256, (h, w)).astype(np.uint16)
        rr, ra = ref9.dwt_forward(img, st, filt)
        orr, oa = oracle9.dwt_forward(img, st, filt)
        assert rr == orr
        if rr == 0:
            assert (ra == oa).all()
            rr2, rb = ref9.dwt_inverse(ra, st, filt)
            or2, ob = oracle9.dwt_inverse(oa, st, filt)
            assert rr2 == or2 and (rb == ob).all()


def test_random_small_images_vs_ref(ref9, oracle9):
    rng = np.random.default_rng(7)
    for it in range(60):
        w, h = int(rng.integers(5, 90)), int(rng.integers(5, 90))
        stages = int(rng.integers(1, 5))
        if -(-w // (1 << stages)) < 3 or -(-h // (1 << stages)) < 3:
            continue
        segments = int(rng.integers(1, 12))
        hw, hh = (-(-w // (1 << (stages - 1)))) // 2, (-(-h // (1 << (stages - 1)))) // 2
        if segments > hw * hh:      # reference behaviour undefined there (icer_compress.c:401)
            continue
        filt = int(rng.integers(0, 7))
        chan = 3 if it % 4 == 0 else 1
        kind = it % 3
        if kind == 0:
            imgs = [rng.integers(0, 256, (h, w)).astype(np.uint16) for _ in range(chan)]
        elif kind == 1:
            base = np.add.outer(np.arange(h) * 2, np.arange(w) * 3) % 200
            imgs = [(base + rng.integers(0, 6, (h, w))).astype(np.uint16) for _ in range(chan)]
        else:
            imgs = [rng.integers(0, 4096, (h, w)).astype(np.uint16) for _ in range(chan)]
        quota = int(rng.choice([40, 500, 3000, w * h * chan, 4 * w * h * chan]))
        a = ref9.compress(imgs, stages, filt, segments, quota)
        b = oracle9.compress(imgs, stages, filt, segments, quota)
        key = (w, h, stages, segments, filt, chan, quota)
        assert a["rc"] == b["rc"], key
        assert a["size_used"] == b["size_used"], key
        assert a["stream"] == b["stream"] and a["prio"] == b["prio"], key
        if a["rc"] in (0, -5):
            assert all((x == y).all() for x, y in zip(a["coefs"], b["coefs"])), key
            da = ref9.decompress(a["stream"], chan, w * h, stages, filt, segments)
            db = oracle9.decompress(a["stream"], chan, w * h, stages, filt, segments)
            if a["size_used"] > 0:
                assert da["rc"] == db["rc"], key
                if da["images"] is not None:
                    assert all((x == y).all() for x, y in zip(da["images"], db["images"])), key


def test_corrupted_and_truncated_streams(ref9, oracle9):
    img = harness.load_boat()[:128, :160].copy()
    enc = ref9.compress([img], 3, 0, 5, 128 * 160)
    s = bytearray(enc["stream"])
    rng = np.random.default_rng(3)
    variants = [bytes(s[: len(s) // 2]), bytes(s[: len(s) - 7]), bytes(s[100:])]
    for _ in range(6):
        t = bytearray(s)
        for pos in rng.integers(0, len(t), 20):
            t[pos] ^= int(rng.integers(1, 256))
        variants.append(bytes(t))
    for v in variants:
        da = ref9.decompress(v, 1, 128 * 160, 3, 0, 5)
        db = oracle9.decompress(v, 1, 128 * 160, 3, 0, 5)
        assert da["rc"] == db["rc"]
        assert (da["images"][0] == db["images"][0]).all()


def test_15_plane_build_vs_ref(ref15, oracle15):
    fr = oracle15.synth(200, 136, 2, 5, 3)   # 12-bit noise
    a = ref15.compress([fr], 4, 0, 10, 2 * fr.size)
    b = oracle15.compress([fr], 4, 0, 10, 2 * fr.size)
    assert a["rc"] == b["rc"] and a["stream"] == b["stream"]
    da = ref15.decompress(a["stream"], 1, fr.size, 4, 0, 10)
    db = oracle15.decompress(a["stream"], 1, fr.size, 4, 0, 10)
    assert (da["images"][0] == db["images"][0]).all()


@pytest.mark.skipif(not harness.have_color_ref(), reason="oracle/_ref/libcolor_ref.so not built")
def test_color_restatement_vs_compiled_macros():
    """harness.rgb_to_yuv (the numpy restatement the fixtures were made with) against the reference's own macros, compiled
    (example/inc/color_util.h:27-29 behind oracle/color_ref.c) -- over ALL 2^24 RGB triples."""
    cr = harness.ColorRef()
    g = np.arange(256, dtype=np.uint8)
    for r in range(0, 256, 1):
        rgb = np.stack(np.broadcast_arrays(np.uint8(r), g[:, None], g[None, :]), axis=-1)
        want = cr.rgb_to_yuv(rgb)
        got = harness.rgb_to_yuv(rgb)
        assert all((a == b).all() for a, b in zip(want, got)), r
