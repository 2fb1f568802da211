#!/bin/bash
# source-level ncu captures of the two coder kernels + the all-kernel DRAM pass (config 2)
OUT=gpurun_out
mkdir -p $OUT
for k in k_encode k_decode; do
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:^$k -c 1 -f -o $OUT/r2_s7_$k \
        python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu --no-e2e --watchdog 350 > $OUT/r2_s7_$k.log 2>&1
    ncu -i $OUT/r2_s7_$k.ncu-rep --page raw --csv > $OUT/r2_s7_${k}_raw.csv 2>/dev/null
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum \
    --clock-control none -c 120 --csv --page raw --log-file $OUT/r2_s7_allkernels_raw.csv \
    python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu --no-e2e --watchdog 500 > $OUT/r2_s7_allkernels.log 2>&1
echo "ncu rc=$?"
