#!/usr/bin/env python
"""Record what the REFERENCE's own programs (built against lib_icer: oracle/_ref/tools_ref, `make -C oracle tools_ref`)
write for the fixture images -> tests/golden/tools_golden.json.  tests/test_tools_dropin.py runs the same programs built
against libicer_b200.so and compares.  Run where /root/reference exists."""
import hashlib
import json
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
TOOLS = os.path.join(ROOT, "oracle", "_ref", "tools_ref")
REF = "/root/reference"


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def main():
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for f in ("boat.512.bmp", "boatcolor.512.bmp"):
            shutil.copy(os.path.join(REF, f), d)
        run = lambda *a: subprocess.run(list(a), cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        u = os.path.join(TOOLS, "icer_util")
        cases = {
            "util_gray_s6_g10_A": ([u, "compress", "boat.512.bmp", "g.icer", "-s", "6", "-g", "10", "-G"],
                                   [u, "decompress", "g.icer", "g.bmp", "-s", "6", "-g", "10", "-G"], "g.icer", "g.bmp"),
            "util_gray_default": ([u, "compress", "boat.512.bmp", "gd.icer", "-G"], [u, "decompress", "gd.icer", "gd.bmp", "-G"], "gd.icer", "gd.bmp"),
            "util_gray_s5_g7_F_t60000": ([u, "compress", "boat.512.bmp", "g2.icer", "-s", "5", "-g", "7", "-f", "F", "-t", "60000", "-G"],
                                         [u, "decompress", "g2.icer", "g2.bmp", "-s", "5", "-g", "7", "-f", "F", "-G"], "g2.icer", "g2.bmp"),
            "util_color_g10_t70000": ([u, "compress", "boatcolor.512.bmp", "c.icer", "-c", "-g", "10", "-t", "70000"],
                                      [u, "decompress", "c.icer", "c.bmp", "-c", "-g", "10"], "c.icer", "c.bmp"),
        }
        for name, (enc, dec, s, b) in cases.items():
            run(*enc); run(*dec)
            out[name] = dict(enc=[os.path.basename(x) if x == u else x for x in enc], dec=[os.path.basename(x) if x == u else x for x in dec],
                             stream=s, image=b, stream_size=os.path.getsize(os.path.join(d, s)), stream_md5=md5(os.path.join(d, s)),
                             image_md5=md5(os.path.join(d, b)))
        # the fixed-parameter demo programs (example_encode.c etc.)
        run(os.path.join(TOOLS, "compress")); run(os.path.join(TOOLS, "decompress"))
        out["example_gray"] = dict(stream="compressed.bin", image="decompress.bmp", stream_size=os.path.getsize(os.path.join(d, "compressed.bin")),
                                   stream_md5=md5(os.path.join(d, "compressed.bin")), image_md5=md5(os.path.join(d, "decompress.bmp")))
        for f in ("compressed.bin", "decompress.bmp"):
            os.remove(os.path.join(d, f))
        run(os.path.join(TOOLS, "compress_color")); run(os.path.join(TOOLS, "decompress_color"))
        names = [f for f in os.listdir(d) if f.endswith(".bin") or (f.endswith(".bmp") and f not in ("boat.512.bmp", "boatcolor.512.bmp", "g.bmp", "gd.bmp", "g2.bmp", "c.bmp"))]
        out["example_color"] = {f: dict(size=os.path.getsize(os.path.join(d, f)), md5=md5(os.path.join(d, f))) for f in sorted(names)}
    json.dump(out, open(os.path.join(HERE, "tools_golden.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
