#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
for v in main pump1; do
  if [ $v = main ]; then unset ICER_B200_LIB; else export ICER_B200_LIB=$PWD/icer_compression_b200/lib/variants/libicer_b200_$v.so; fi
  for wl in c2 c4; do
    B=""; [ $wl = c4 ] && B="--batch 64"
    timeout 200 python bench.py --workload $wl $B --no-cpu --no-e2e --steps 3 --warmup 2 --watchdog 180 > $OUT/var_${v}_$wl.json 2>$OUT/var_${v}_$wl.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/var_${v}_$wl.json"))
    r=d["roofline"]; o=d["roofline_other"]
    ke = r["kernel_ms"] if r["kernel"].startswith("k_encode") else o["kernel_ms"]
    kd = r["kernel_ms"] if r["kernel"].startswith("k_decode") else o["kernel_ms"]
    print("$v $wl", "value", d["value"], "enc", d["encode_ms"], "dec", d["decode_ms"], "k_encode", ke, "k_decode", kd)
except Exception as e: print("$v $wl failed", e)
PY
  done
done
unset ICER_B200_LIB
timeout 500 ncu --set full --clock-control none --import-source on -k regex:^k_encode -c 1 -f -o $OUT/r2_s11_c4_k_encode \
    python bench.py --workload c4 --batch 32 --steps 1 --warmup 0 --no-cpu --no-e2e --watchdog 450 > $OUT/r2_s11_c4_k_encode.log 2>&1
ncu -i $OUT/r2_s11_c4_k_encode.ncu-rep --page raw --csv > $OUT/r2_s11_c4_k_encode_raw.csv 2>/dev/null
echo "ncu rc=$?"
