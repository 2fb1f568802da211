"""GPU parity tests for lib_icer's NATIVE uint8 codec: icer_compress_image_uint8 / _yuv_uint8, the
decoders and the uint8 wavelet transform -- int8 arithmetic, 7 bit planes, 300 packets, the ascending colour segment order
(A.16) and the odd-length interleave defect (A.5) -- against committed golden vectors generated from the unmodified
reference and, when oracle/_ref travelled to the box, against the reference itself.  The bar is equality."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import harness

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD8 = json.load(open(os.path.join(HERE, "golden", "golden_u8.json")))["cases"]


def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()


@pytest.fixture(scope="module")
def ctx():
    import icer_compression_b200 as icer
    c = icer.IcerB200(device=0, planes=9, max_stages=6)
    yield c
    c.close()


@pytest.fixture(scope="module")
def legacy():
    import icer_compression_b200 as icer
    return icer.LegacyAPI()


@pytest.mark.parametrize("c", GOLD8, ids=[c["name"] for c in GOLD8])
def test_u8_golden_legacy_api(c, legacy, oracle9):
    from make_golden_u8 import case_input_u8
    imgs = case_input_u8(c, oracle9)
    r = legacy.compress(imgs, c["stages"], c["filt"], c["segments"], c["quota"], dtype=np.uint8)
    assert r["rc"] == c["rc"]
    assert r["size_used"] == c["size_used"]
    if c["size_used"]:
        assert md5(r["stream"]) == c["stream_md5"]
        assert md5(r["prio"]) == c["prio_md5"]
        assert md5(b"".join(p.tobytes() for p in r["coefs"])) == c["coefs_md5"]      # what lib_icer leaves in the image buffers
        d = legacy.decompress(r["stream"], c["chan"], c["w"] * c["h"], c["stages"], c["filt"], c["segments"], dtype=np.uint8)
        assert d["rc"] == c["dec_rc"] and (d["w"], d["h"]) == (c["w"], c["h"])
        assert md5(b"".join(p.tobytes() for p in d["images"])) == c["dec_md5"]


def test_u8_golden_batch_api(ctx, oracle9):
    """The same vectors through icer_b200_*_batch_pix(ICER_B200_PIX_NATIVE_U8), several frames per call."""
    import icer_compression_b200 as icer
    from make_golden_u8 import case_input_u8
    for c in GOLD8:
        if c["rc"] not in (0, harness.QUOTA):
            continue
        imgs = np.stack(case_input_u8(c, oracle9))                      # [chan, h, w]
        frames = np.stack([imgs, imgs[:, ::-1].copy(), imgs])           # frame 1 differs: the batch must not mix frames up
        r = ctx.compress_batch_pix(frames, icer.PIX_NATIVE_U8, c["w"], c["h"], c["chan"], c["stages"], c["filt"], c["segments"], c["quota"])
        for i in (0, 2):
            assert r["rc"][i] == c["rc"] and int(r["size_used"][i]) == c["size_used"]
            assert md5(r["streams"][i]) == c["stream_md5"]
        assert md5(r["streams"][1]) != c["stream_md5"]
        d = ctx.decompress_batch_pix(r["streams"], icer.PIX_NATIVE_U8, c["chan"], c["w"], c["h"], c["stages"], c["filt"], c["segments"])
        for i in (0, 2):
            assert d["rc"][i] == 0 and tuple(d["dims"][i]) == (c["w"], c["h"])
            assert md5(d["frames"][i].tobytes()) == c["dec_md5"]


def test_u8_packet_limit_is_encoder_only(ctx, legacy, oracle9):
    """uint8 colour at 6 stages = 399 packets > ICER_MAX_PACKETS (300): lib_icer's encoder refuses; with the limit raised the
    stream is produced, and decoding never depended on the limit."""
    import icer_compression_b200 as icer
    from make_golden_u8 import case_input_u8
    c = [g for g in GOLD8 if g["name"] == "u8_color_s6_packets"][0]
    imgs = np.stack(case_input_u8(c, oracle9))[None]
    c2 = icer.IcerB200(device=0)
    try:
        r = c2.compress_batch_pix(imgs, icer.PIX_NATIVE_U8, 512, 512, 3, 6, 0, 10, 70000)
        assert r["rc"][0] == -9 and r["size_used"][0] == 0
        c2.set_limits_u8(max_packets=800)
        r = c2.compress_batch_pix(imgs, icer.PIX_NATIVE_U8, 512, 512, 3, 6, 0, 10, 70000)
        assert r["rc"][0] == harness.QUOTA and 0 < r["size_used"][0] <= 70000
        c2.set_limits_u8(max_packets=300)
        d = c2.decompress_batch_pix(r["streams"], icer.PIX_NATIVE_U8, 3, 512, 512, 6, 0, 10)
        assert d["rc"][0] == 0
    finally:
        c2.close()


needs_ref = pytest.mark.skipif(not harness.have_ref(), reason="oracle/_ref not on this box")


@needs_ref
@pytest.mark.parametrize("filt", range(7))
def test_u8_dwt_vs_reference(filt, ctx, ref9):
    """Forward and inverse uint8 transform, even and ODD sizes (the inverse's odd-length placement is lib_icer's defect)."""
    rng = np.random.default_rng(800 + filt)
    for (w, h, st) in ((64, 48, 3), (37, 53, 2), (5, 5, 1), (129, 65, 5), (96, 96, 4), (200, 7, 1), (301, 203, 5), (511, 509, 6)):
        for hi in (64, 128, 256):          # 6-bit (safe), 7-bit, full 8-bit (int8 wrap + overflow flag)
            img = rng.integers(0, hi, (h, w)).astype(np.uint8)
            rc, out = ctx.dwt_batch(img[None], st, filt, dtype=np.uint8)
            rrc, ro = ref9.dwt8(img, st, filt)
            assert rc[0] == rrc, (w, h, st, hi)
            assert (out[0] == ro).all(), (w, h, st, hi)
            _, back = ctx.dwt_batch(ro[None], st, filt, inverse=True, dtype=np.uint8)
            _, rb = ref9.dwt8(ro, st, filt, inverse=True)
            assert (back[0] == rb).all(), (w, h, st, hi, "inverse")


@needs_ref
def test_u8_random_images_vs_reference(ctx, legacy, ref9):
    import icer_compression_b200 as icer
    rng = np.random.default_rng(4242)
    shapes = [(64, 64, 1, 3, 4), (97, 131, 1, 3, 6), (131, 97, 3, 3, 5), (200, 120, 1, 4, 10), (33, 45, 3, 2, 3), (256, 256, 1, 5, 32),
              (61, 200, 1, 3, 1)]
    for k, (w, h, chan, st, seg) in enumerate(shapes):
        filt = k % 7
        bits = (5, 6, 7, 8)[k % 4]             # 8-bit data overflows the int8 transform on most images: rc must agree
        yy, xx = np.mgrid[0:h, 0:w]
        base = ((np.sin(xx / 9.0 + k) + np.cos(yy / 7.0)) * 0.2 + 0.5) * ((1 << bits) - 1)
        imgs = [np.clip(base + rng.integers(-3, 4, (h, w)), 0, (1 << bits) - 1).astype(np.uint8) for _ in range(chan)]
        for quota in (w * h * chan * 2, max(400, w * h * chan // 6)):
            a = ref9.compress(imgs, st, filt, seg, quota, dtype=np.uint8)
            b = legacy.compress(imgs, st, filt, seg, quota, dtype=np.uint8)
            assert a["rc"] == b["rc"], (k, quota)
            assert a["size_used"] == b["size_used"], (k, quota)
            if a["rc"] not in (0, harness.QUOTA):
                continue
            assert a["stream"] == b["stream"] and a["prio"] == b["prio"], (k, quota)
            for p, q in zip(a["coefs"], b["coefs"]):
                assert (p == q).all(), (k, quota, "coefs")
            if a["size_used"]:
                da = ref9.decompress(a["stream"], chan, w * h, st, filt, seg, dtype=np.uint8)
                db = legacy.decompress(a["stream"], chan, w * h, st, filt, seg, dtype=np.uint8)
                assert da["rc"] == db["rc"] and (da["w"], da["h"]) == (db["w"], db["h"])
                for p, q in zip(da["images"], db["images"]):
                    assert (p == q).all(), (k, quota, "decoded")
                # the batch entry point agrees with the legacy wrapper
                dc = ctx.decompress_batch_pix([a["stream"]], icer.PIX_NATIVE_U8, chan, w, h, st, filt, seg)
                assert (dc["frames"][0].reshape(chan, h, w) == np.stack(da["images"])).all()


@needs_ref
def test_u8_truncated_stream_vs_reference(legacy, ref9, oracle9):
    from make_golden_u8 import case_input_u8
    c = [g for g in GOLD8 if g["name"] == "u8_synth_odd"][0]
    imgs = case_input_u8(c, oracle9)
    s = ref9.compress(imgs, c["stages"], c["filt"], c["segments"], c["quota"], dtype=np.uint8)["stream"]
    for cut in (len(s) // 3, len(s) // 2, len(s) - 5):
        bad = bytearray(s[:cut])
        bad[len(bad) // 2] ^= 0x40
        da = ref9.decompress(bytes(bad), 1, c["w"] * c["h"], c["stages"], c["filt"], c["segments"], dtype=np.uint8)
        db = legacy.decompress(bytes(bad), 1, c["w"] * c["h"], c["stages"], c["filt"], c["segments"], dtype=np.uint8)
        assert da["rc"] == db["rc"]
        assert (da["images"][0] == db["images"][0]).all(), cut
