"""GPU tests of opt-in, not yet measured code paths.  They sort last and are non-blocking (xfail, strict=False): an XPASS in
the log means the path works and can be measured; nothing else depends on them."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx9():
    import icer_compression_b200 as icer
    c = icer.IcerB200(device=0, planes=9, max_stages=6)
    yield c
    c.close()


@pytest.mark.xfail(reason="ICER_B200_BULK_COPY is an opt-in host path written after the last GPU session of round 1; "
                          "it has not run on a GPU yet", strict=False)
def test_bulk_copy_host_path(ctx9, oracle9):
    import icer_compression_b200 as icer
    frames = np.stack([oracle9.synth(160, 96, f, 5, 0) for f in range(40)])[:, None]
    ref = ctx9.compress_batch(frames, 4, 0, 10, 160 * 96)
    os.environ["ICER_B200_BULK_COPY"] = "1"
    try:
        c = icer.IcerB200(device=0, planes=9, max_stages=6)
    finally:
        del os.environ["ICER_B200_BULK_COPY"]
    try:
        enc = c.compress_batch(frames, 4, 0, 10, 160 * 96)
        assert (enc["rc"] == ref["rc"]).all() and enc["streams"] == ref["streams"]
        dec = c.decompress_batch(enc["streams"], 1, 160, 96, 4, 0, 10)
        assert (dec["rc"] == 0).all() and (dec["frames"].reshape(frames.shape) == frames).all()
    finally:
        c.close()
