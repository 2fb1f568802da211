#!/bin/bash
# Round-2 GPU session 3: first run of the bit-sliced coder kernels -- smoke, test-suite, quick timing of c2/c3/c5.
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python __graft_entry__.py smoke > $OUT/r2_s3_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/r2_s3_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -rfEx --timeout 300 -x > $OUT/r2_s3_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r2_s3_pytest.log
tail -15 $OUT/r2_s3_pytest.log
for wl in c2 c3 c5; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --steps 5 --warmup 3 --watchdog 250 > $OUT/r2_s3_$wl.json 2>$OUT/r2_s3_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r2_s3_$wl.json"))
    r=d["roofline"]
    print("$wl", "value", d["value"], "enc", d["encode_ms"], "dec", d["decode_ms"], "lossless", d["roundtrip_lossless"], r["kernel"][:8], r["kernel_ms"], d["roofline_other"]["kernel_ms"], "dwt", r["encode_dwt_ms"], r["decode_dwt_ms"])
except Exception as e: print("$wl failed", e)
PY
done
