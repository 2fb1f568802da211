"""profiles/r2_sign_stats.py -- CPU only: how often does SOME of the 32 frames of a group need a sign decision at a pixel (lockstep pays a\nsecond pass then), how many steps does per-lane stepping with a sync every 32 pixels need, and what is the ideal (one frame alone)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import harness
orc = harness.Oracle(planes=9, max_stages=6)
n, w, h = 32, 512, 512
mags = np.stack([(orc.compress([orc.synth(w, h, f, 20260924, 0)], 6, 0, 10, w * h)["coefs"][0] & 0x7FFF).astype(np.int32) for f in range(n)])
print(mags.shape)
tot_cur = tot_el = tot_ideal = 0.0
for p in range(8, -1, -1):
    first = ((mags >> p) == 1)           # becomes significant at plane p -> sign decision
    q = first.mean()
    anysign = first.any(axis=0).mean()
    # chunks of 32 px along rows
    fc = first.reshape(n, h, w // 32, 32).sum(axis=3)     # signs per chunk per lane
    el = (32 + fc.max(axis=0)).mean() / 32               # elastic: iterations per pixel (ignoring quiet)
    ideal = 1 + q
    print("plane %d: P(sign)=%.3f  P(any of 32 sign)=%.3f  passes/pixel: lockstep %.2f elastic32 %.2f ideal %.2f" % (p, q, anysign, 1 + anysign, el, ideal))
