#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference (run where /root/reference exists).

  python tests/golden/make_golden.py

Writes (all small):
  boat512.u8.gz            512x512 gray of /root/reference/boat.512.bmp (stb's 3->1 channel conversion,
                           example/inc/stb_image.h:1747-1750), the input of BASELINE config 1
  boatcolor512.rgb.u8.gz   RGB pixels of boatcolor.512.bmp (tests derive Y,Cb,Cr with the tools' integer conversion,
                           example/inc/color_util.h:27-29, example/src/icer_util.c:69-92)
  golden.json              for a list of (input, params): rc, size_used, md5 of the rearranged stream,
                           md5 of the decoded image(s), md5 of the coefficient array the reference leaves
                           in the caller's buffer -- all produced by oracle/_ref/libicer_ref{,15}.so
  streams/*.bin.gz         a few full compressed streams (decode-side fixtures)
"""
import gzip
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness  # noqa: E402

REF = "/root/reference"


def read_bmp24(path):
    raw = open(path, "rb").read()
    off = int.from_bytes(raw[10:14], "little")
    w = int.from_bytes(raw[18:22], "little", signed=True)
    h = int.from_bytes(raw[22:26], "little", signed=True)
    bpp = int.from_bytes(raw[28:30], "little")
    assert bpp == 24 and w > 0
    stride = (w * 3 + 3) & ~3
    rows = np.frombuffer(raw, dtype=np.uint8, count=stride * abs(h), offset=off).reshape(abs(h), stride)
    bgr = rows[:, :w * 3].reshape(abs(h), w, 3)
    if h > 0:
        bgr = bgr[::-1]
    return bgr[:, :, ::-1].astype(np.int64)  # RGB


def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()


def main():
    rgb = read_bmp24(os.path.join(REF, "boat.512.bmp"))
    gray = ((rgb[:, :, 0] * 77 + rgb[:, :, 1] * 150 + rgb[:, :, 2] * 29) >> 8).astype(np.uint8)
    with gzip.GzipFile(os.path.join(HERE, "boat512.u8.gz"), "wb", mtime=0) as f:
        f.write(gray.tobytes())

    rgb = read_bmp24(os.path.join(REF, "boatcolor.512.bmp"))
    r, g, b = rgb[:, :, 0], rgb[:, :, 1], rgb[:, :, 2]
    clip = lambda x: np.clip(x, 0, 255)
    y = clip((19595 * r + 38470 * g + 7471 * b) >> 16)
    cb = clip(((36962 * (b - y)) >> 16) + 128)
    cr = clip(((46727 * (r - y)) >> 16) + 128)
    yuv = np.stack([y, cb, cr]).astype(np.uint8)
    with gzip.GzipFile(os.path.join(HERE, "boatcolor512.rgb.u8.gz"), "wb", mtime=0) as f:
        f.write(rgb.astype(np.uint8).tobytes())

    ref9 = harness.Ref(False)
    ref15 = harness.Ref(True)
    orc = harness.Oracle()  # only for the synthetic generator
    boat = gray.astype(np.uint16)
    col = [p.astype(np.uint16) for p in yuv]

    cases = []

    def add(name, ref, planes_in, stages, filt, segments, quota, keep_stream=False):
        enc = ref.compress(planes_in, stages, filt, segments, quota)
        h, w = planes_in[0].shape
        dec = ref.decompress(enc["stream"], len(planes_in), w * h, stages, filt, segments)
        rec = dict(name=name, planes=ref.planes, chan=len(planes_in), w=w, h=h, stages=stages, filt=filt,
                   segments=segments, quota=quota, rc=enc["rc"], size_used=enc["size_used"],
                   stream_md5=md5(enc["stream"]), prio_md5=md5(enc["prio"]),
                   coefs_md5=md5(b"".join(c.tobytes() for c in enc["coefs"])),
                   dec_rc=dec["rc"],
                   decoded_md5=md5(b"".join(i.tobytes() for i in dec["images"])) if dec["images"] else None)
        if keep_stream:
            os.makedirs(os.path.join(HERE, "streams"), exist_ok=True)
            with gzip.GzipFile(os.path.join(HERE, "streams", name + ".bin.gz"), "wb", mtime=0) as f:
                f.write(enc["stream"])
        cases.append(rec)
        print(name, enc["rc"], enc["size_used"], rec["stream_md5"])

    # BASELINE config 1 and the per-filter KATs of SURVEY.md 8(c)
    for fname, f in harness.FILTERS.items():
        add("boat_s6_g10_%s" % fname, ref9, [boat], 6, f, 10, 262144, keep_stream=(fname == "A"))
    add("boat_s4_g6_A_q30000", ref9, [boat], 4, 0, 6, 30000, keep_stream=True)       # example_encode.c defaults
    add("boat_s4_g10_A_q70000", ref9, [boat], 4, 0, 10, 70000)
    add("boatcolor_s4_g10_A_q70000", ref9, col, 4, 0, 10, 70000, keep_stream=True)   # BASELINE config 3 on the real image
    add("boatcolor_s6_g10_A_q70000", ref9, col, 6, 0, 10, 70000)
    add("boatcolor_s4_g6_A_q100000", ref9, col, 4, 0, 6, 100000)                     # example_encode_color.c defaults
    # synthetic frames (generator of SURVEY.md 8(d), oracle/icer_oracle.c:orc_synth_frame)
    for kind, kn in ((0, "nat8"), (1, "noise8")):
        fr = orc.synth(512, 512, 3, 20260922 + 2, kind)
        add("synth512_%s_s6_g10_A" % kn, ref9, [fr], 6, 0, 10, 512 * 512)
        add("synth512_%s_s6_g10_A_q2x" % kn, ref9, [fr], 6, 0, 10, 2 * 512 * 512)
    yuvs = [orc.synth(512, 512, f, 20260922 + 3, 0) for f in (5, 1005, 2005)]
    add("synth512_yuv_s4_g10_A_q70000", ref9, yuvs, 4, 0, 10, 70000)
    fr = orc.synth(320, 200, 1, 20260922 + 4, 2)
    add("synth320x200_nat12_s5_g10_A_15planes", ref15, [fr], 5, 0, 10, 2 * 320 * 200)
    fr = orc.synth(1024, 1024, 0, 20260922 + 4, 2)
    add("synth1024_nat12_s7_g10_A_15planes", ref15, [fr], 7, 0, 10, 2 * 1024 * 1024)
    fr = orc.synth(301, 203, 7, 99, 0)
    for fname, f in harness.FILTERS.items():
        add("synth301x203_nat8_s5_g7_%s" % fname, ref9, [fr], 5, f, 7, 301 * 203)
    fr = orc.synth(1024, 1024, 11, 20260922 + 5, 0)
    add("synth1024_nat8_s6_g10_A", ref9, [fr], 6, 0, 10, 1024 * 1024)

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_golden.py", reference="TheRealOrange/icer_compression@0e2961a",
                       cases=cases), f, indent=1)


if __name__ == "__main__":
    main()
