for v in main base; do
  if [ $v = main ]; then unset ICER_B200_LIB; else export ICER_B200_LIB=$PWD/icer_compression_b200/lib/variants/libicer_b200_$v.so; fi
  timeout 100 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 --watchdog 80 > gpurun_out/var_$v.json 2>gpurun_out/var_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_$v.json"))
    print("$v", "value", d["value"], "enc", d["encode_ms"], "dec", d["decode_ms"], "lossless", d["roundtrip_lossless"], d["roofline"]["kernel"][:8], d["roofline"]["kernel_ms"], d["roofline_other"]["kernel_ms"])
except Exception as e: print("$v failed", e)
PY
done
