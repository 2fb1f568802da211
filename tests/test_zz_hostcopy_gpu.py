"""GPU tests of the host-copy variants (sorted last)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_bulk_copy_host_path(oracle9):
    """The host paths move every slot's used bytes with ONE strided copy per direction (default); ICER_B200_BULK_COPY=0 is the
    copy-per-frame fallback.  Both must give the same bytes."""
    import icer_compression_b200 as icer
    frames = np.stack([oracle9.synth(160, 96, f, 5, 0) for f in range(40)])[:, None]
    # 1170 segments of 28 header bytes do not fit the 15360-byte quota: every frame ends with ICER_BYTE_QUOTA_EXCEEDED and
    # streams of different lengths -- what the strided copies have to get right
    os.environ["ICER_B200_BULK_COPY"] = "0"
    try:
        c0 = icer.IcerB200(device=0, planes=9, max_stages=6)
    finally:
        del os.environ["ICER_B200_BULK_COPY"]
    ref = c0.compress_batch(frames, 4, 0, 10, 160 * 96)
    for i in range(40):
        assert ref["streams"][i] == oracle9.compress([frames[i, 0]], 4, 0, 10, 160 * 96)["stream"], i
    c0.close()
    os.environ["ICER_B200_BULK_COPY"] = "1"
    try:
        c = icer.IcerB200(device=0, planes=9, max_stages=6)
    finally:
        del os.environ["ICER_B200_BULK_COPY"]
    try:
        enc = c.compress_batch(frames, 4, 0, 10, 160 * 96)
        assert (enc["rc"] == ref["rc"]).all() and enc["streams"] == ref["streams"]
        dec = c.decompress_batch(enc["streams"], 1, 160, 96, 4, 0, 10)
        assert (dec["rc"] == 0).all()
        for i in range(40):
            o = oracle9.decompress(enc["streams"][i], 1, 160 * 96, 4, 0, 10)
            assert (dec["frames"][i, 0] == o["images"][0].ravel()).all(), i
    finally:
        c.close()
