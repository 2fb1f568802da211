#!/bin/bash
# Round-2 opening experiments (build here, run the loop of run_variants.sh on a GPU box).  Rationale: a dense bit plane of a
# level-1 segment takes ~30 ms of a warp's time at full occupancy (~5 us per pixel), yet blocked warps re-poll every
# 0.1-0.3 us; ~20 % of k_decode's issued instructions are such polls (profiles/r1_v8_k_decode_lines.txt).
set -e
cd "$(dirname "$0")/.."
profiles/variants.sh poll5us   "-DICER_DEC_BACKOFF_NS=5000  -DICER_DEC_SLOTWAIT_NS=20000 -DICER_DEC_CHAINWAIT_NS=5000"
profiles/variants.sh poll20us  "-DICER_DEC_BACKOFF_NS=20000 -DICER_DEC_SLOTWAIT_NS=50000 -DICER_DEC_CHAINWAIT_NS=20000"
profiles/variants.sh poll20r64 "-DICER_DEC_BACKOFF_NS=20000 -DICER_DEC_SLOTWAIT_NS=50000 -DICER_DEC_CHAINWAIT_NS=20000 -DICER_DEC_RECHECK=64"
profiles/variants.sh poll1us   "-DICER_DEC_BACKOFF_NS=1000  -DICER_DEC_SLOTWAIT_NS=5000  -DICER_DEC_CHAINWAIT_NS=1000"
echo "then on the GPU box:  for v in main poll1us poll5us poll20us poll20r64 -> bash profiles/run_variants.sh (edit its list)"
