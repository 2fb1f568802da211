"""Drop-in proof: the reference's OWN programs (example/src/icer_util.c and the four demo programs),
compiled from the reference tree without modification against include/icer.h and linked against libicer_b200.so
(oracle/Makefile target `tools_b200` -> oracle/_ref/tools_b200/), must write exactly what the same programs write when
built against lib_icer (tests/golden/tools_golden.json, recorded by tests/golden/make_tools_golden.py)."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import harness

HERE = os.path.dirname(os.path.abspath(__file__))
TOOLS = os.path.join(harness.ROOT, "oracle", "_ref", "tools_b200")
GOLD = json.load(open(os.path.join(HERE, "golden", "tools_golden.json")))


def have_tools():
    return os.path.exists(os.path.join(TOOLS, "icer_util"))


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def test_tools_link_only_the_seven_entry_points():
    """CPU: the tools resolve against libicer_b200.so and import from it exactly what they import from lib_icer."""
    if not have_tools():
        pytest.skip("oracle/_ref/tools_b200 not built")
    ldd = subprocess.check_output(["ldd", os.path.join(TOOLS, "icer_util")]).decode()
    assert "libicer_b200.so" in ldd and "not found" not in ldd
    want = {"icer_init", "icer_init_output_struct", "icer_compress_image_uint16", "icer_compress_image_yuv_uint16",
            "icer_decompress_image_uint16", "icer_decompress_image_yuv_uint16", "icer_get_image_dimensions"}
    got = set()
    for t in ("icer_util", "compress", "decompress", "compress_color", "decompress_color"):
        syms = subprocess.check_output(["nm", "-u", os.path.join(TOOLS, t)]).decode()
        got |= {ln.split()[-1].split("@")[0] for ln in syms.splitlines() if " icer_" in ln}
    assert got == want


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("tools")
    boat = harness.load_boat().astype(np.uint8)
    harness.write_bmp24(str(d / "boat.512.bmp"), np.repeat(boat[:, :, None], 3, axis=2))
    harness.write_bmp24(str(d / "boatcolor.512.bmp"), harness.load_boatcolor_rgb())
    return d


@pytest.mark.gpu
@pytest.mark.parametrize("name", [k for k in GOLD if k.startswith("util_")])
def test_icer_util_cli(name, workdir):
    if not have_tools():
        pytest.skip("oracle/_ref/tools_b200 not built")
    g = GOLD[name]
    u = os.path.join(TOOLS, "icer_util")
    for cmd in (g["enc"], g["dec"]):
        subprocess.run([u] + cmd[1:], cwd=str(workdir), check=True, stdout=subprocess.DEVNULL, timeout=300)
    assert os.path.getsize(workdir / g["stream"]) == g["stream_size"]
    assert md5(workdir / g["stream"]) == g["stream_md5"]
    assert md5(workdir / g["image"]) == g["image_md5"]


@pytest.mark.gpu
def test_example_programs(workdir):
    if not have_tools():
        pytest.skip("oracle/_ref/tools_b200 not built")
    run = lambda t: subprocess.run([os.path.join(TOOLS, t)], cwd=str(workdir), check=True, stdout=subprocess.DEVNULL, timeout=300)
    run("compress"); run("decompress")
    g = GOLD["example_gray"]
    assert md5(workdir / g["stream"]) == g["stream_md5"] and md5(workdir / g["image"]) == g["image_md5"]
    os.remove(workdir / "compressed.bin"); os.remove(workdir / "decompress.bmp")
    run("compress_color"); run("decompress_color")
    for f, rec in GOLD["example_color"].items():
        assert os.path.getsize(workdir / f) == rec["size"] and md5(workdir / f) == rec["md5"], f
