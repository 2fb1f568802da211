#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
for wl in c2 c4 c5 c2noise; do
    timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --steps 3 --warmup 2 --watchdog 280 > $OUT/r2_s12_$wl.json 2>$OUT/r2_s12_$wl.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/r2_s12_$wl.json"))
    r=d["roofline"]; o=d["roofline_other"]
    ke = r["kernel_ms"] if r["kernel"].startswith("k_encode") else o["kernel_ms"]
    kd = r["kernel_ms"] if r["kernel"].startswith("k_decode") else o["kernel_ms"]
    print("$wl", "value", d["value"], "enc", d["encode_ms"], "dec", d["decode_ms"], "k_encode", ke, "k_decode", kd, "lossless", d["roundtrip_lossless"])
except Exception as e: print("$wl failed", e)
PY
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_atsize.py -m gpu -q -x --timeout 300 2>&1 | tail -2
