#!/bin/bash
# Round-2 GPU session 2: test-suite on the new kernels (fused DWT stages both ways, vector layout kernels, shared-memory codeword
# window, priority-prefix encoding), quick timing, source-level ncu captures of the two coder kernels, all-kernel DRAM pass.
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -rfEx --timeout 600 -x > $OUT/r2_s2_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r2_s2_pytest.log
tail -4 $OUT/r2_s2_pytest.log
for wl in c2 c3 c5; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --steps 5 --warmup 3 --watchdog 250 > $OUT/r2_s2_$wl.json 2>$OUT/r2_s2_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r2_s2_$wl.json"))
    r=d["roofline"]
    print("$wl", "value", d["value"], "enc", d["encode_ms"], "dec", d["decode_ms"], "lossless", d["roundtrip_lossless"], r["kernel"][:8], r["kernel_ms"], d["roofline_other"]["kernel_ms"], "dwt", r["encode_dwt_ms"], r["decode_dwt_ms"])
except Exception as e: print("$wl failed", e)
PY
done
for k in k_encode k_decode; do
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:^$k -c 1 -f -o $OUT/r2_s2_$k \
        python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu --no-e2e --watchdog 350 > $OUT/r2_s2_$k.log 2>&1
    ncu -i $OUT/r2_s2_$k.ncu-rep --page raw --csv > $OUT/r2_s2_${k}_raw.csv 2>/dev/null
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum \
    --clock-control none -c 120 --csv --page raw --log-file $OUT/r2_s2_allkernels_raw.csv \
    python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu --no-e2e --watchdog 500 > $OUT/r2_s2_allkernels.log 2>&1
echo "ncu rc=$?"
