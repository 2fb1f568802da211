#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rfEx --timeout 300 -x > $OUT/r2_s5_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r2_s5_pytest.log
tail -4 $OUT/r2_s5_pytest.log
for wl in c2 c2noise c3 c5; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --steps 5 --warmup 3 --watchdog 250 > $OUT/r2_s5_$wl.json 2>$OUT/r2_s5_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r2_s5_$wl.json"))
    r=d["roofline"]
    print("$wl", "value", d["value"], "enc", d["encode_ms"], "dec", d["decode_ms"], "lossless", d["roundtrip_lossless"], r["kernel"][:8], r["kernel_ms"], d["roofline_other"]["kernel_ms"], "dwt", r["encode_dwt_ms"], r["decode_dwt_ms"])
except Exception as e: print("$wl failed", e)
PY
done
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > $OUT/r2_s5_sanitizer.log 2>&1; tail -3 $OUT/r2_s5_sanitizer.log
