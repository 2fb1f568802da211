#!/bin/bash
# vectorised k_scan, parallel header build in k_gather: test-suite, config 2 timing, kernel list
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rfEx --timeout 300 -x > $OUT/r2_s18_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r2_s18_pytest.log
tail -4 $OUT/r2_s18_pytest.log
for wl in c2 c3; do
timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --steps 5 --warmup 3 --watchdog 280 > $OUT/r2_s18_$wl.json 2>$OUT/r2_s18_$wl.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/r2_s18_$wl.json"))
    r=d["roofline"]; o=d["roofline_other"]
    ke = r["kernel_ms"] if r["kernel"].startswith("k_encode") else o["kernel_ms"]
    kd = r["kernel_ms"] if r["kernel"].startswith("k_decode") else o["kernel_ms"]
    print("$wl", "value", d["value"], "enc", d["encode_ms"], "dec", d["decode_ms"], "k_encode", ke, "k_decode", kd, "dwt", r["encode_dwt_ms"], r["decode_dwt_ms"], "lossless", d["roundtrip_lossless"])
except Exception as e: print("$wl failed", e)
PY
done
bash profiles/r2_kernlist.sh r2_s18 c2
