#!/bin/bash
# bench each library build named on the command line (main = the in-tree library) on config 2, coder kernels only
mkdir -p gpurun_out
for v in "$@"; do
  if [ $v = main ]; then unset ICER_B200_LIB; else export ICER_B200_LIB=$PWD/icer_compression_b200/lib/variants/libicer_b200_$v.so; fi
  timeout 100 python bench.py --workload c2 --no-cpu --no-e2e --steps 5 --warmup 3 --watchdog 80 > gpurun_out/var_$v.json 2>gpurun_out/var_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_$v.json"))
    r=d["roofline"]; o=d["roofline_other"]
    ke = r["kernel_ms"] if r["kernel"].startswith("k_encode") else o["kernel_ms"]
    kd = r["kernel_ms"] if r["kernel"].startswith("k_decode") else o["kernel_ms"]
    print("$v", "value", d["value"], "enc", d["encode_ms"], "dec", d["decode_ms"], "lossless", d["roundtrip_lossless"], "k_encode", ke, "k_decode", kd)
except Exception as e: print("$v failed", e)
PY
done
