#!/bin/bash
# Round-2 GPU session 1: host facts, the whole GPU test-suite (incl. the new at-size parity tests), the full bench line,
# the decoder polling variants and an all-kernel ncu pass (duration + DRAM bytes per launch of every kernel of one step).
OUT=gpurun_out
mkdir -p $OUT
{ nproc; free -g; nvidia-smi -L; nvidia-smi topo -m; } > $OUT/r2_s1_host.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -rfEx --timeout 600 > $OUT/r2_s1_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r2_s1_pytest.log
tail -5 $OUT/r2_s1_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r2_s1_bench.json 2> $OUT/r2_s1_bench.err
echo "bench rc=$?"; tail -c 600 $OUT/r2_s1_bench.err
for v in main poll1us poll5us poll20us poll20r64; do
  if [ $v = main ]; then unset ICER_B200_LIB; else export ICER_B200_LIB=$PWD/icer_compression_b200/lib/variants/libicer_b200_$v.so; fi
  timeout 120 python bench.py --workload c2 --no-cpu --no-e2e --steps 5 --warmup 3 --watchdog 100 > $OUT/r2_var_$v.json 2>$OUT/r2_var_$v.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r2_var_$v.json"))
    print("$v", "value", d["value"], "enc", d["encode_ms"], "dec", d["decode_ms"], "lossless", d["roundtrip_lossless"], d["roofline"]["kernel"][:8], d["roofline"]["kernel_ms"], d["roofline_other"]["kernel_ms"])
except Exception as e: print("$v failed", e)
PY
done
unset ICER_B200_LIB
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats \
    --clock-control none -c 120 --csv --page raw --log-file $OUT/r2_s1_allkernels_raw.csv \
    python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu --no-e2e --watchdog 500 > $OUT/r2_s1_allkernels.log 2>&1
echo "ncu rc=$?"
