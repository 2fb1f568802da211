#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rfEx --timeout 300 -x > $OUT/r2_s15_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r2_s15_pytest.log
tail -4 $OUT/r2_s15_pytest.log
for v in fused split; do
    if [ $v = split ]; then export ICER_B200_SPLIT_SCAN=1; else unset ICER_B200_SPLIT_SCAN; fi
    timeout 300 python bench.py --workload c2 --no-cpu --no-e2e --steps 5 --warmup 3 --watchdog 280 > $OUT/r2_s15_$v.json 2>$OUT/r2_s15_$v.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/r2_s15_$v.json"))
    r=d["roofline"]; o=d["roofline_other"]
    ke = r["kernel_ms"] if r["kernel"].startswith("k_encode") else o["kernel_ms"]
    kd = r["kernel_ms"] if r["kernel"].startswith("k_decode") else o["kernel_ms"]
    print("$v", "value", d["value"], "enc", d["encode_ms"], "dec", d["decode_ms"], "k_encode", ke, "k_decode", kd, "dwt", r["encode_dwt_ms"], r["decode_dwt_ms"], "lossless", d["roundtrip_lossless"])
except Exception as e: print("$v failed", e)
PY
done
