#!/bin/bash
# prefetch / poll-loop variants on config 2, prefetch on a small batch of 4096^2 frames (latency-bound)
OUT=gpurun_out
mkdir -p $OUT
bash profiles/run_variants2.sh main nopf poll1 poll2
for v in main nopf; do
  if [ $v = main ]; then unset ICER_B200_LIB; else export ICER_B200_LIB=$PWD/icer_compression_b200/lib/variants/libicer_b200_$v.so; fi
  timeout 300 python bench.py --workload c4 --batch 32 --no-cpu --no-e2e --steps 2 --warmup 1 --watchdog 280 > $OUT/var_${v}_c4.json 2>$OUT/var_${v}_c4.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/var_${v}_c4.json"))
    print("$v c4@32", "value", d["value"], "enc", d["encode_ms"], "dec", d["decode_ms"])
except Exception as e: print("$v c4 failed", e)
PY
done
