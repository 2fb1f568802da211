#!/usr/bin/env python
"""profiles/lockstep_stats.py -- how much does "lane = frame in lockstep" cost on the benchmark frames?  CPU only (uses the
oracle to get the coefficients).  For each bit plane: the fraction of pixels at which ONE frame needs the slow path (some
pixel of its 3x3 neighbourhood is significant at that plane) versus the fraction at which ANY of the 32 frames of a group
does -- the warp pays the second number.  Result for config 2's generator (32 frames): slow decisions per pixel summed over
the 9 planes 3.51 (one frame) vs 4.15 (any of 32): lockstep costs 1.18x, so letting lanes drift apart can gain at most 18 %;
the 16-19 active lanes per instruction come from divergence INSIDE the slow path (bin kinds, codeword completions, sign
decisions), not from quiet lanes waiting."""
import os
import sys

import numpy as np
from scipy.ndimage import maximum_filter

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import harness  # noqa: E402

orc = harness.Oracle(planes=9, max_stages=6)
n, w, h = 32, 512, 512
mags = np.stack([(orc.compress([orc.synth(w, h, f, 20260924, 0)], 6, 0, 10, w * h)["coefs"][0] & 0x7FFF).astype(np.int32) for f in range(n)])
nb = np.stack([maximum_filter(m, size=3, mode="constant") for m in mags])    # ignores segment borders: small error
print("plane  slow(one frame)  slow(any of 32)  uncoded(cat 3)")
s1 = s32 = 0.0
for p in range(8, -1, -1):
    nq = (nb >> p) != 0
    one, anyl = nq.mean(), nq.any(axis=0).mean()
    s1 += one; s32 += anyl
    print("%5d  %15.3f  %15.3f  %14.3f" % (p, one, anyl, ((mags >> p) >= 8).mean()))
print("sum over planes: %.3f vs %.3f -> lockstep overhead %.2fx" % (s1, s32, s32 / s1))
