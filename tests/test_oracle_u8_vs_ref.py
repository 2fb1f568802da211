"""CPU: the oracle's restatement of lib_icer's NATIVE uint8 codec against the unmodified reference (oracle/_ref) and the
committed golden vectors (tests/golden/golden_u8.json).  Pins the oracle for SURVEY 8(f) rank 3: int8 wavelet arithmetic,
7 planes / 300 packets, ascending colour segment order (A.16), the odd-length interleave defect (A.5)."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import harness

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD8 = json.load(open(os.path.join(HERE, "golden", "golden_u8.json")))["cases"]


def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()


@pytest.mark.parametrize("c", GOLD8, ids=[c["name"] for c in GOLD8])
def test_oracle_u8_matches_golden(c, oracle9):
    from make_golden_u8 import case_input_u8
    imgs = case_input_u8(c, oracle9)
    r = oracle9.compress(imgs, c["stages"], c["filt"], c["segments"], c["quota"], dtype=np.uint8)
    assert r["rc"] == c["rc"] and r["size_used"] == c["size_used"]
    if not c["size_used"]:
        return
    assert md5(r["stream"]) == c["stream_md5"] and md5(r["prio"]) == c["prio_md5"]
    assert md5(b"".join(p.tobytes() for p in r["coefs"])) == c["coefs_md5"]
    d = oracle9.decompress(r["stream"], c["chan"], c["w"] * c["h"], c["stages"], c["filt"], c["segments"], dtype=np.uint8)
    assert d["rc"] == c["dec_rc"] and (d["w"], d["h"]) == (c["w"], c["h"])
    assert md5(b"".join(p.tobytes() for p in d["images"])) == c["dec_md5"]


def test_oracle_interleave8_map_vs_ref(oracle9, ref9):
    for n in list(range(2, 600)) + [1023, 1024, 1025, 3001]:
        assert (oracle9.interleave8_map(n) == ref9.interleave8_map(n)).all(), n


@pytest.mark.parametrize("filt", range(7))
def test_oracle_u8_dwt_vs_ref(filt, oracle9, ref9):
    rng = np.random.default_rng(900 + filt)
    for (w, h, st) in ((64, 48, 3), (37, 53, 2), (5, 5, 1), (129, 65, 5), (200, 7, 1), (301, 203, 5)):
        for hi in (64, 128, 256):
            img = rng.integers(0, hi, (h, w)).astype(np.uint8)
            a_rc, a = oracle9.dwt8(img, st, filt)
            b_rc, b = ref9.dwt8(img, st, filt)
            assert a_rc == b_rc and (a == b).all(), (w, h, st, hi)
            _, ai = oracle9.dwt8(b, st, filt, inverse=True)
            _, bi = ref9.dwt8(b, st, filt, inverse=True)
            assert (ai == bi).all(), (w, h, st, hi, "inverse")


def test_oracle_u8_codec_vs_ref(oracle9, ref9):
    rng = np.random.default_rng(31337)
    shapes = [(64, 64, 1, 3, 4), (97, 131, 1, 3, 6), (131, 97, 3, 3, 5), (200, 120, 1, 4, 10), (33, 45, 3, 2, 3), (61, 200, 1, 3, 1),
              (128, 128, 3, 4, 8)]
    for k, (w, h, chan, st, seg) in enumerate(shapes):
        filt = k % 7
        bits = (5, 6, 7, 8)[k % 4]
        yy, xx = np.mgrid[0:h, 0:w]
        base = ((np.sin(xx / 9.0 + k) + np.cos(yy / 7.0)) * 0.2 + 0.5) * ((1 << bits) - 1)
        imgs = [np.clip(base + rng.integers(-3, 4, (h, w)), 0, (1 << bits) - 1).astype(np.uint8) for _ in range(chan)]
        for quota in (w * h * chan * 2, max(400, w * h * chan // 6)):
            a = ref9.compress(imgs, st, filt, seg, quota, dtype=np.uint8)
            b = oracle9.compress(imgs, st, filt, seg, quota, dtype=np.uint8)
            assert a["rc"] == b["rc"] and a["size_used"] == b["size_used"], (k, quota)
            if a["rc"] not in (0, harness.QUOTA) or not a["size_used"]:
                continue
            assert a["stream"] == b["stream"] and a["prio"] == b["prio"], (k, quota)
            for p, q in zip(a["coefs"], b["coefs"]):
                assert (p == q).all(), (k, quota)
            da = ref9.decompress(a["stream"], chan, w * h, st, filt, seg, dtype=np.uint8)
            db = oracle9.decompress(a["stream"], chan, w * h, st, filt, seg, dtype=np.uint8)
            assert da["rc"] == db["rc"] and (da["w"], da["h"]) == (db["w"], db["h"])
            for p, q in zip(da["images"], db["images"]):
                assert (p == q).all(), (k, quota, "decoded")
            # truncated + bit-flipped stream: same reconstruction
            bad = bytearray(a["stream"][: max(60, len(a["stream"]) * 2 // 3)])
            bad[len(bad) // 2] ^= 0x10
            da = ref9.decompress(bytes(bad), chan, w * h, st, filt, seg, dtype=np.uint8)
            db = oracle9.decompress(bytes(bad), chan, w * h, st, filt, seg, dtype=np.uint8)
            assert da["rc"] == db["rc"]
            if da["images"] is not None:
                for p, q in zip(da["images"], db["images"]):
                    assert (p == q).all(), (k, quota, "damaged")
