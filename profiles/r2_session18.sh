#!/bin/bash
# A/B of library builds (every variant compiled from source by profiles/variants.sh): unified magnitude / sign set-up, neighbour-bin
# probe, forced flush in the same step
OUT=gpurun_out
mkdir -p $OUT
bash profiles/run_variants2.sh prev unified probe iflush prev unified
for v in unified iflush; do
  export ICER_B200_LIB=$PWD/icer_compression_b200/lib/variants/libicer_b200_$v.so
  timeout 300 python bench.py --workload c5 --no-cpu --no-e2e --steps 3 --warmup 2 --watchdog 280 > $OUT/var_${v}_c5.json 2>$OUT/var_${v}_c5.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/var_${v}_c5.json"))
    print("$v c5", "value", d["value"], "enc", d["encode_ms"], "dec", d["decode_ms"])
except Exception as e: print("$v c5 failed", e)
PY
done
