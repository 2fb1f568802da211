#!/bin/bash
# profiles/capture.sh TAG -- the ncu passes behind profiles/<TAG>_*: run on a GPU box from the repo root.
# 1. launch list of the default bench command (per-launch times, cold cache, serialised -> kernel SHARES)
# 2. one --set full capture of each of the three dominant kernels
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --watchdog 150 > $OUT/${TAG}_launches.log 2>&1
for k in k_encode k_decode k_fwd_stage_tma; do
    timeout 240 ncu --set full --clock-control none --import-source on -k regex:^$k -c 1 -f -o $OUT/${TAG}_$k \
        python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --watchdog 200 > $OUT/${TAG}_$k.log 2>&1
    ncu -i $OUT/${TAG}_$k.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_raw.csv 2>/dev/null
done
