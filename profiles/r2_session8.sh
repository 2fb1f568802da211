#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rfEx --timeout 300 -x > $OUT/r2_s8_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r2_s8_pytest.log
tail -4 $OUT/r2_s8_pytest.log
for w in 2 3 4; do
  timeout 300 python bench.py --workload c2 --no-cpu --steps 5 --warmup 3 --e2e-workers $w --watchdog 250 > $OUT/r2_s8_e2e$w.json 2>$OUT/r2_s8_e2e$w.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r2_s8_e2e$w.json"))
    r=d["roofline"]
    print("workers $w", "value", d["value"], "enc", d["encode_ms"], "dec", d["decode_ms"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "dwt", r["encode_dwt_ms"], r["decode_dwt_ms"])
except Exception as e: print("$w failed", e)
PY
done
