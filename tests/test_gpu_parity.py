"""GPU parity tests: the CUDA path, called through the C ABI, must be bit-exact against the CPU oracle
(oracle/liboracle.so), the committed golden vectors (generated from the unmodified reference) and -- when the
prebuilt oracle/_ref/*.so travelled to the box -- the reference itself.  Integer work: the bar is equality."""
import hashlib
import json
import os

import numpy as np
import pytest

import harness

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))["cases"]


def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()


@pytest.fixture(scope="module")
def ctx9():
    import icer_compression_b200 as icer
    c = icer.IcerB200(device=0, planes=9, max_stages=6)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx15():
    import icer_compression_b200 as icer
    c = icer.IcerB200(device=0, planes=15, max_stages=7)
    yield c
    c.close()


@pytest.fixture(scope="module")
def legacy():
    import icer_compression_b200 as icer
    return icer.LegacyAPI()


def case_input(c, orc):
    from test_oracle_vs_ref import case_input as ci
    return ci(c, orc)


# ---- wavelet -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("filt", range(7))
def test_dwt_forward_inverse_vs_oracle(filt, ctx9, oracle9):
    rng = np.random.default_rng(200 + filt)
    # widths that are multiples of 8 take the fused TMA stage kernels (odd region sizes appear at their later stages: 200 -> 25,
    # 120 -> 15, 56 -> 7); the others the two-pass kernels
    for (w, h, st) in ((64, 48, 3), (37, 53, 2), (5, 5, 1), (129, 65, 5), (96, 96, 5), (200, 7, 1), (512, 512, 6), (301, 203, 5),
                       (200, 56, 4), (120, 56, 4), (128, 70, 3), (256, 7, 1), (72, 40, 3), (1024, 64, 4)):
        imgs = np.stack([rng.integers(0, 256, (h, w)).astype(np.uint16),
                         rng.integers(0, 4096, (h, w)).astype(np.uint16),
                         rng.integers(0, 65536, (h, w)).astype(np.uint16)])
        rc, out = ctx9.dwt_batch(imgs, st, filt)
        for i in range(3):
            orc_rc, o = oracle9.dwt_forward(imgs[i], st, filt)
            assert rc[i] == orc_rc, (w, h, st, i)
            assert (out[i] == o).all(), (w, h, st, i)
        rc2, back = ctx9.dwt_batch(out, st, filt, inverse=True)
        for i in range(3):
            _, o = oracle9.dwt_inverse(out[i], st, filt)
            assert (back[i] == o).all(), (w, h, st, i, "inverse")


def test_dwt_too_many_stages(ctx9, legacy):
    img = np.zeros((1, 16, 16), dtype=np.uint16)
    rc, _ = ctx9.dwt_batch(img, 4, 0)
    assert rc[0] == harness.TOO_MANY_STAGES
    assert legacy.dwt_forward(img[0], 4, 0)[0] == harness.TOO_MANY_STAGES


# ---- golden vectors through the batch API and the legacy API ------------------------------------------------
@pytest.mark.parametrize("c", GOLD, ids=[c["name"] for c in GOLD])
def test_golden_encode_decode(c, ctx9, ctx15, oracle9, oracle15):
    ctx, orc = (ctx15, oracle15) if c["planes"] == 15 else (ctx9, oracle9)
    imgs = case_input(c, orc)
    frames = np.stack(imgs)[None]
    enc = ctx.compress_batch(frames, c["stages"], c["filt"], c["segments"], c["quota"], want_prio=True, want_coefs=True)
    assert int(enc["rc"][0]) == c["rc"]
    assert int(enc["size_used"][0]) == c["size_used"]
    assert md5(enc["streams"][0]) == c["stream_md5"]
    assert md5(enc["prio"][0]) == c["prio_md5"]
    assert md5(enc["coefs"][0].tobytes()) == c["coefs_md5"]
    dec = ctx.decompress_batch(enc["streams"], c["chan"], c["w"], c["h"], c["stages"], c["filt"], c["segments"])
    assert int(dec["rc"][0]) == c["dec_rc"]
    assert tuple(dec["dims"][0]) == (c["w"], c["h"])
    assert md5(dec["frames"][0].tobytes()) == c["decoded_md5"]


def test_golden_streams_decode_only(ctx9):
    """Streams written by the REFERENCE (committed fixtures) decode to the reference's pixels."""
    import gzip
    for c in GOLD:
        p = os.path.join(HERE, "golden", "streams", c["name"] + ".bin.gz")
        if not os.path.exists(p):
            continue
        s = gzip.open(p, "rb").read()
        assert md5(s) == c["stream_md5"]
        dec = ctx9.decompress_batch([s], c["chan"], c["w"], c["h"], c["stages"], c["filt"], c["segments"])
        assert int(dec["rc"][0]) == 0 and md5(dec["frames"][0].tobytes()) == c["decoded_md5"]


def test_legacy_api_boat(legacy, oracle9):
    """BASELINE config 1 through the drop-in entry points (icer_compress_image_uint16 / icer_decompress_image_uint16)."""
    boat = harness.load_boat()
    g = [c for c in GOLD if c["name"] == "boat_s6_g10_A"][0]
    a = legacy.compress([boat], 6, 0, 10, 262144)
    assert a["rc"] == 0 and a["size_used"] == g["size_used"] and md5(a["stream"]) == g["stream_md5"]
    assert md5(a["prio"]) == g["prio_md5"] and md5(a["coefs"][0].tobytes()) == g["coefs_md5"]
    d = legacy.decompress(a["stream"], 1, 512 * 512, 6, 0, 10)
    assert d["rc"] == 0 and (d["w"], d["h"]) == (512, 512) and md5(d["images"][0].tobytes()) == g["decoded_md5"]
    assert (d["images"][0] == boat).all()     # lossless at this quota
    # too-small output buffer: ICER_BYTE_QUOTA_EXCEEDED like icer_compress.c:461-463
    assert legacy.decompress(a["stream"], 1, 1000, 6, 0, 10)["rc"] == harness.QUOTA


def test_legacy_api_color(legacy):
    yuv = harness.load_boatcolor_yuv()
    g = [c for c in GOLD if c["name"] == "boatcolor_s4_g10_A_q70000"][0]
    a = legacy.compress(yuv, 4, 0, 10, 70000)
    assert a["rc"] == harness.QUOTA and a["size_used"] == g["size_used"] and md5(a["stream"]) == g["stream_md5"]
    d = legacy.decompress(a["stream"], 3, 512 * 512, 4, 0, 10)
    assert d["rc"] == 0 and md5(b"".join(i.tobytes() for i in d["images"])) == g["decoded_md5"]


# ---- randomised shapes / parameters vs the oracle -----------------------------------------------------------
def test_random_images_vs_oracle(ctx9, oracle9):
    rng = np.random.default_rng(11)
    done = 0
    for it in range(200):
        w, h = int(rng.integers(5, 140)), int(rng.integers(5, 140))
        stages = int(rng.integers(1, 6))
        if -(-w // (1 << stages)) < 3 or -(-h // (1 << stages)) < 3:
            continue
        segments = int(rng.integers(1, 33))
        hw, hh = (-(-w // (1 << (stages - 1)))) // 2, (-(-h // (1 << (stages - 1)))) // 2
        if segments > hw * hh:
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
        key = (w, h, stages, segments, filt, chan, quota)
        a = oracle9.compress(imgs, stages, filt, segments, quota)
        b = ctx9.compress_batch(np.stack(imgs)[None], stages, filt, segments, quota, want_prio=True, want_coefs=True)
        assert a["rc"] == int(b["rc"][0]), key
        assert a["size_used"] == int(b["size_used"][0]), key
        assert a["stream"] == b["streams"][0] and a["prio"] == b["prio"][0], key
        if a["rc"] in (0, -5):
            assert all((x == y).all() for x, y in zip(a["coefs"], b["coefs"][0])), key
            if a["size_used"] > 0:
                da = oracle9.decompress(a["stream"], chan, w * h, stages, filt, segments)
                db = ctx9.decompress_batch([a["stream"]], chan, w, h, stages, filt, segments)
                assert da["rc"] == int(db["rc"][0]), key
                assert all((x.ravel() == y).all() for x, y in zip(da["images"], db["frames"][0])), key
        done += 1
    assert done > 60


def test_batch_of_frames_vs_oracle(ctx9, oracle9):
    """A batch of different frames in one call: every stream equals the oracle's single-image result."""
    n = 12
    frames = np.stack([oracle9.synth(256, 256, f, 20260922 + 2, f % 2) for f in range(n)])[:, None]
    quota = 256 * 256
    enc = ctx9.compress_batch(frames, 5, 0, 10, quota)
    for i in range(n):
        o = oracle9.compress([frames[i, 0]], 5, 0, 10, quota)
        assert o["rc"] == int(enc["rc"][i]) and o["stream"] == enc["streams"][i], i
    dec = ctx9.decompress_batch(enc["streams"], 1, 256, 256, 5, 0, 10)
    for i in range(n):
        o = oracle9.decompress(enc["streams"][i], 1, 256 * 256, 5, 0, 10)
        assert int(dec["rc"][i]) == 0 and (dec["frames"][i, 0] == o["images"][0].ravel()).all(), i


def test_scratch_retry_path(oracle9):
    """Force the payload scratch to be too small: the encoder must retry transparently and still be bit-exact."""
    import icer_compression_b200 as icer
    c = icer.IcerB200(device=0)
    c.set_scratch_bits(1)
    fr = oracle9.synth(128, 128, 1, 7, 1)   # noise: > 1 bit / pixel / plane in the low planes
    enc = c.compress_batch(fr[None, None], 3, 0, 6, 2 * 128 * 128)
    o = oracle9.compress([fr], 3, 0, 6, 2 * 128 * 128)
    assert o["stream"] == enc["streams"][0]
    c.close()


def test_corrupted_and_truncated_streams(ctx9, oracle9):
    img = harness.load_boat()[:128, :160].copy()
    s = bytearray(oracle9.compress([img], 3, 0, 5, 128 * 160)["stream"])
    rng = np.random.default_rng(3)
    variants = [bytes(s[: len(s) // 2]), bytes(s[: len(s) - 7]), bytes(s[100:])]
    for _ in range(6):
        t = bytearray(s)
        for pos in rng.integers(0, len(t), 20):
            t[pos] ^= int(rng.integers(1, 256))
        variants.append(bytes(t))
    dec = ctx9.decompress_batch(variants, 1, 160, 128, 3, 0, 5)
    for i, v in enumerate(variants):
        o = oracle9.decompress(v, 1, 128 * 160, 3, 0, 5)
        assert o["rc"] == int(dec["rc"][i])
        assert (o["images"][0].ravel() == dec["frames"][i, 0]).all(), i


def test_error_codes(ctx9):
    fr = np.full((1, 1, 64, 64), 65535, dtype=np.uint16)         # wavelet overflow -> ICER_INTEGER_OVERFLOW, nothing written
    fr[0, 0, ::2, ::2] = 0
    enc = ctx9.compress_batch(fr, 2, 0, 4, 64 * 64)
    assert int(enc["rc"][0]) == harness.OVERFLOW and int(enc["size_used"][0]) == 0
    enc = ctx9.compress_batch(np.zeros((1, 1, 16, 16), dtype=np.uint16), 4, 0, 2, 1000)
    assert int(enc["rc"][0]) == harness.TOO_MANY_STAGES
    enc = ctx9.compress_batch(np.zeros((1, 1, 16, 16), dtype=np.uint16), 2, 0, 33, 1000)
    assert int(enc["rc"][0]) == harness.TOO_MANY_SEGMENTS


def test_vs_reference_when_present(ctx9):
    """If oracle/_ref/libicer_ref.so travelled to this box, compare against the reference itself."""
    if not harness.have_ref():
        pytest.skip("oracle/_ref not present")
    ref = harness.Ref(False)
    orc = harness.Oracle()
    fr = orc.synth(512, 512, 9, 20260922 + 2, 0)
    a = ref.compress([fr], 6, 0, 10, 512 * 512)
    b = ctx9.compress_batch(fr[None, None], 6, 0, 10, 512 * 512)
    assert a["rc"] == int(b["rc"][0]) and a["stream"] == b["streams"][0]
    da = ref.decompress(a["stream"], 1, 512 * 512, 6, 0, 10)
    db = ctx9.decompress_batch([a["stream"]], 1, 512, 512, 6, 0, 10)
    assert (da["images"][0].ravel() == db["frames"][0, 0]).all()


def test_device_synth_generator_matches_oracle(ctx9, oracle9):
    """bench.py fills HBM with icer_b200_synth_frames_u16; it must equal the generator the CPU legs use."""
    import torch
    for kind, (w, h) in ((0, (96, 64)), (1, (64, 48)), (2, (80, 40)), (3, (33, 17))):
        t = torch.empty((3, h, w), dtype=torch.int16, device="cuda:0")
        ctx9.synth_frames(t, 3, w, h, 5, 1234, kind)
        torch.cuda.synchronize()
        got = t.cpu().numpy().view(np.uint16)
        for i in range(3):
            assert (got[i] == oracle9.synth(w, h, 5 + i, 1234, kind)).all(), (kind, i)


def test_device_pointer_api_roundtrip(ctx9, oracle9):
    """The batch calls with DEVICE pointers (what bench.py's `value` times) agree with the host-pointer calls."""
    import torch
    import icer_compression_b200 as icer
    n, w, h = 6, 128, 96
    fr = np.stack([oracle9.synth(w, h, f, 77, 0) for f in range(n)])[:, None]
    host = ctx9.compress_batch(fr, 3, 0, 7, w * h)
    d_fr = torch.from_numpy(fr.view(np.int16)).cuda()
    out = torch.zeros((n, w * h), dtype=torch.uint8, device="cuda")
    used = torch.zeros(n, dtype=torch.int64, device="cuda")
    rc = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx9.compress_batch_raw((w, h, 1, 3, 0, 7, w * h), d_fr, n, out, w * h, None, None, used, rc, icer.MEM_DEVICE)
    torch.cuda.synchronize()
    for i in range(n):
        assert bytes(out[i, :int(used[i])].cpu().numpy()) == host["streams"][i]
    dec = torch.zeros((n, 1, w * h), dtype=torch.int16, device="cuda")
    dims = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
    drc = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx9.decompress_batch_raw((w, h, 1, 3, 0, 7, 0), out, w * h, used, n, dec, dims, drc, icer.MEM_DEVICE)
    torch.cuda.synchronize()
    assert (drc.cpu().numpy() == 0).all()
    hostdec = ctx9.decompress_batch(host["streams"], 1, w, h, 3, 0, 7)
    assert (dec.cpu().numpy().view(np.uint16) == hostdec["frames"]).all()


def _segments(stream):
    """(offset, length) of every segment of a well-formed stream."""
    out, off = [], 0
    while off + 28 <= len(stream):
        nbits = int.from_bytes(stream[off + 16:off + 20], "little")
        ln = 28 + (nbits + 7) // 8
        out.append((off, ln))
        off += ln
    return out


def _recrc(seg):
    import zlib
    seg = bytearray(seg)
    seg[20:24] = zlib.crc32(bytes(seg[28:])).to_bytes(4, "little")
    seg[24:28] = zlib.crc32(bytes(seg[:24])).to_bytes(4, "little")
    return bytes(seg)


def test_crc_valid_garbage_payloads(ctx9, oracle9):
    """Segments whose payload is garbage but whose CRCs are valid: the entropy decoder runs into invalid codewords /
    out-of-data mid-plane; lib_icer keeps what that plane wrote so far and never decodes the planes below it
    (icer_partition.c:441).  The pipelined GPU decoder has to reproduce exactly that."""
    rng = np.random.default_rng(21)
    img = harness.load_boat()[64:64 + 96, 100:100 + 128].copy()
    base = oracle9.compress([img], 2, 0, 4, 4 * 96 * 128)["stream"]
    segs = _segments(base)
    variants = []
    for trial in range(24):
        s = bytearray(base)
        for _ in range(int(rng.integers(1, 4))):
            off, ln = segs[int(rng.integers(0, len(segs)))]
            if ln <= 30:
                continue
            seg = bytearray(s[off:off + ln])
            mode = trial % 3
            if mode == 0:      # random payload bytes
                seg[28:] = rng.integers(0, 256, ln - 28, dtype=np.uint8).tobytes()
            elif mode == 1:    # claim fewer bits than there are (truncation inside the payload)
                nbits = int.from_bytes(seg[16:20], "little")
                nb2 = int(rng.integers(1, max(2, nbits // 2)))
                seg[16:20] = nb2.to_bytes(4, "little")
                seg = seg[:28 + (nb2 + 7) // 8]
                s[off:off + ln] = _recrc(seg) + b"\0" * (ln - len(seg))
                continue
            else:              # all ones / all zeros
                seg[28:] = (b"\xff" if trial % 2 else b"\x00") * (ln - 28)
            s[off:off + ln] = _recrc(seg)
        variants.append(bytes(s))
    dec = ctx9.decompress_batch(variants, 1, 128, 96, 2, 0, 4)
    for i, v in enumerate(variants):
        o = oracle9.decompress(v, 1, 96 * 128, 2, 0, 4)
        assert o["rc"] == int(dec["rc"][i]), i
        assert (o["images"][0].ravel() == dec["frames"][i, 0]).all(), i


def test_scanner_overflow_fallback(ctx9, oracle9):
    """More segment candidates than the device scanner holds -> that frame is scanned on the host; same result."""
    img = harness.load_boat()[:128, :160].copy()
    s = oracle9.compress([img], 3, 0, 5, 128 * 160)["stream"]
    ref = ctx9.decompress_batch([s, s[: len(s) // 2]], 1, 160, 128, 3, 0, 5)
    os.environ["ICER_B200_PARSE_CAP"] = "16"
    try:
        got = ctx9.decompress_batch([s, s[: len(s) // 2]], 1, 160, 128, 3, 0, 5)
    finally:
        del os.environ["ICER_B200_PARSE_CAP"]
    assert (ref["rc"] == got["rc"]).all() and (ref["frames"] == got["frames"]).all()
    o = oracle9.decompress(s, 1, 128 * 160, 3, 0, 5)
    assert (o["images"][0].ravel() == got["frames"][0, 0]).all()


def test_mixed_dimensions_in_one_batch(ctx9, oracle9):
    """Streams of different image sizes in one decode call: dimensions come from the headers, like lib_icer."""
    a = oracle9.synth(96, 64, 1, 5, 0)
    b = oracle9.synth(64, 80, 2, 5, 1)
    sa = oracle9.compress([a], 3, 0, 4, 96 * 64)["stream"]
    sb = oracle9.compress([b], 3, 0, 4, 64 * 80)["stream"]
    dec = ctx9.decompress_batch([sa, sb, sa, b"", sb], 1, 100, 80, 3, 0, 4)     # capacity 8000 px per slot
    assert list(dec["rc"]) == [0, 0, 0, -7, 0]
    assert [tuple(d) for d in dec["dims"]] == [(96, 64), (64, 80), (96, 64), (0, 0), (64, 80)]
    oa = oracle9.decompress(sa, 1, 8000, 3, 0, 4)["images"][0].ravel()
    ob = oracle9.decompress(sb, 1, 8000, 3, 0, 4)["images"][0].ravel()
    assert (dec["frames"][0, 0][:oa.size] == oa).all() and (dec["frames"][2, 0][:oa.size] == oa).all()
    assert (dec["frames"][1, 0][:ob.size] == ob).all() and (dec["frames"][4, 0][:ob.size] == ob).all()


def test_pixel_formats_at_the_boundary(ctx9, oracle9):
    """8-bit widening and RGB888 <-> Y,Cb,Cr on the device equal the tools' host-side steps
    (example/src/icer_util.c:69-121,163-168,322-325) around the uint16 codec."""
    import icer_compression_b200 as icer
    rng = np.random.default_rng(5)
    # packed RGB -> YUV -> stream, and back
    rgb = harness.load_boatcolor_rgb()[100:100 + 96, 200:200 + 128].copy()
    noise = rng.integers(0, 256, rgb.shape, dtype=np.uint8)
    batch = np.stack([rgb, noise])
    quota = 3 * 96 * 128
    enc = ctx9.compress_batch_pix(batch, icer.PIX_RGB888, 128, 96, 3, 3, 0, 6, quota)
    for i in range(2):
        yuv = harness.rgb_to_yuv(batch[i])
        o = oracle9.compress(yuv, 3, 0, 6, quota)
        assert o["rc"] == int(enc["rc"][i]) and o["stream"] == enc["streams"][i], i
    dec = ctx9.decompress_batch_pix(enc["streams"], icer.PIX_RGB888, 3, 128, 96, 3, 0, 6)
    for i in range(2):
        y, cb, cr = [p.astype(np.int64).ravel() for p in oracle9.decompress(enc["streams"][i], 3, 96 * 128, 3, 0, 6)["images"]]
        clip = lambda v: np.clip(v, 0, 255)
        want = np.stack([clip(y + ((91881 * cr) >> 16) - 179), clip(y - ((22544 * cb + 46793 * cr) >> 16) + 135),
                         clip(y + ((116129 * cb) >> 16) - 226)], axis=1).astype(np.uint8)
        assert int(dec["rc"][i]) == 0 and (dec["frames"][i] == want).all(), i
    # 8-bit gray: widened on the device; decode clamps to 255
    g8 = np.stack([harness.load_boat()[:96, :128].astype(np.uint8), rng.integers(0, 256, (96, 128), dtype=np.uint8)])[:, None]
    enc = ctx9.compress_batch_pix(g8, icer.PIX_U8, 128, 96, 1, 3, 0, 6, 2 * 96 * 128)
    dec = ctx9.decompress_batch_pix(enc["streams"], icer.PIX_U8, 1, 128, 96, 3, 0, 6)
    for i in range(2):
        o = oracle9.compress([g8[i, 0].astype(np.uint16)], 3, 0, 6, 2 * 96 * 128)
        assert o["stream"] == enc["streams"][i]
        od = oracle9.decompress(o["stream"], 1, 96 * 128, 3, 0, 6)["images"][0].ravel()
        assert (dec["frames"][i, 0] == np.minimum(od, 255).astype(np.uint8)).all()


def test_device_color_kernels_vs_compiled_reference_macros(ctx15):
    """k_rgb_to_yuv / k_yuv_to_rgb against the reference tools' own macros, compiled (example/inc/color_util.h:27-34 behind
    oracle/color_ref.c -> oracle/_ref/libcolor_ref.so): every one of the 2^24 RGB triples forward, and a dense sweep of (Y, Cb, Cr)
    back -- the conversions themselves, not a stream round trip."""
    if not harness.have_color_ref():
        pytest.skip("oracle/_ref/libcolor_ref.so not present")
    cr = harness.ColorRef()
    g = np.arange(256, dtype=np.uint8)
    rgb = np.stack(np.broadcast_arrays(g[:, None, None], g[None, :, None], g[None, None, :]), axis=-1).reshape(4096, 4096, 3).copy()
    want = cr.rgb_to_yuv(rgb)
    # 15 bit planes: lossless for any 8-bit input.  The conversion kernels sit behind the *_pix batch calls; drive them through a 1-stage lossless round trip of the planes:
    # encode RGB888 -> decode as planar uint16 gives exactly the Y, Cb, Cr planes the device produced
    quota = 3 * 4096 * 4096 * 2
    enc = ctx15.compress_batch_pix(rgb[None], 2, 4096, 4096, 3, 1, 0, 32, quota)          # PIX_RGB888
    assert int(enc["rc"][0]) == 0
    dec = ctx15.decompress_batch(enc["streams"], 3, 4096, 4096, 1, 0, 32)
    for c in range(3):
        assert (dec["frames"][0, c].reshape(4096, 4096) == want[c]).all(), c
    # back: planes (Y, Cb, Cr) in 0..255 (what an 8-bit colour image decodes to) -> packed RGB on the device
    yy, cb, crr = [a.reshape(4096, 4096).astype(np.uint16) for a in np.meshgrid(g, g, g, indexing="ij")]
    planes = np.stack([yy, cb, crr])[None]
    enc2 = ctx15.compress_batch(planes, 1, 0, 32, quota)
    assert int(enc2["rc"][0]) == 0
    got = ctx15.decompress_batch_pix(enc2["streams"], 2, 3, 4096, 4096, 1, 0, 32)["frames"][0].reshape(4096, 4096, 3)
    assert (got == cr.yuv_to_rgb(yy, cb, crr)).all()


def test_priority_prefix_encoding(oracle9):
    """Byte quotas far below the lossless size: the encoder codes a prefix of the priority order first and the rest only if some
    frame keeps the whole prefix.  Streams, sizes and status codes must not depend on the prefix length -- forced here to several
    values (ICER_B200_FORCE_PREFIX), on a batch whose frames hit the quota at very different places -- nor on the learned hint."""
    import icer_compression_b200 as icer
    w, h, n = 160, 128, 36
    frames = np.stack([oracle9.synth(w, h, f, 99, (f % 3 == 2) * 1) if f % 5 else np.full((h, w), 7, dtype=np.uint16) for f in range(n)])[:, None]
    for quota in (900, 6000, 40000):
        want = [oracle9.compress([frames[i, 0]], 4, 0, 8, quota) for i in range(n)]
        for prefix in ("1", "37", "300", "600", "935", None):
            if prefix is None:
                os.environ.pop("ICER_B200_FORCE_PREFIX", None)
            else:
                os.environ["ICER_B200_FORCE_PREFIX"] = prefix
            try:
                c = icer.IcerB200(device=0)
            finally:
                os.environ.pop("ICER_B200_FORCE_PREFIX", None)
            for rep in range(3 if prefix is None else 1):          # without forcing: the 2nd and 3rd call use the learned prefix
                enc = c.compress_batch(frames, 4, 0, 8, quota, want_prio=True)
                for i in range(n):
                    assert int(enc["rc"][i]) == want[i]["rc"], (quota, prefix, rep, i)
                    assert enc["streams"][i] == want[i]["stream"] and enc["prio"][i] == want[i]["prio"], (quota, prefix, rep, i)
            c.close()


def test_host_pipeline_chunks(oracle9):
    """Host-buffer batches larger than the pipeline chunk go through overlapped upload / compute / download; results must
    not depend on the chunking."""
    import icer_compression_b200 as icer
    os.environ["ICER_B200_PIPE_CHUNK"] = "5"
    try:
        c = icer.IcerB200(device=0)
    finally:
        del os.environ["ICER_B200_PIPE_CHUNK"]
    n = 13
    frames = np.stack([oracle9.synth(96, 64, f, 11, f % 2) for f in range(n)])[:, None]
    enc = c.compress_batch(frames, 3, 0, 5, 96 * 64)
    for i in range(n):
        o = oracle9.compress([frames[i, 0]], 3, 0, 5, 96 * 64)
        assert o["rc"] == int(enc["rc"][i]) and o["stream"] == enc["streams"][i], i
    dec = c.decompress_batch(enc["streams"], 1, 96, 64, 3, 0, 5)
    for i in range(n):
        o = oracle9.decompress(enc["streams"][i], 1, 96 * 64, 3, 0, 5)
        assert int(dec["rc"][i]) == 0 and (dec["frames"][i, 0] == o["images"][0].ravel()).all(), i
    c.close()


@pytest.mark.timeout(240)
def test_decoder_pool_many_launches(ctx9, oracle9):
    """Regression for a stall of the decoder's task pool: chains used to fetch their global work id after waiting for a
    slot, so a later chain could overtake an earlier one and the end-of-work marker landed on the earlier chain while the
    later one was still valid.  Many launches with many short chains per block (small frames, 32 segments) hit the window."""
    w = h = 96
    n = 96
    frames = np.stack([oracle9.synth(128, 128, f, 77, 0)[:h, :w] for f in range(n)])[:, None]
    enc = ctx9.compress_batch(frames, 3, 0, 32, 200000)      # 2880 segments: 80 kB of headers alone
    assert (enc["rc"] == 0).all()
    ref = ctx9.decompress_batch(enc["streams"], 1, w, h, 3, 0, 32)["frames"]
    assert (ref.reshape(n, 1, h, w) == frames).all()
    for _ in range(60):
        out = ctx9.decompress_batch(enc["streams"], 1, w, h, 3, 0, 32)
        assert (out["rc"] == 0).all()
        assert (out["frames"] == ref).all()
