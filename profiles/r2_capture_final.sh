#!/bin/bash
# Round-2 final captures (config 2): source-level ncu of the two coder kernels, the all-kernel launch list with DRAM bytes.
OUT=gpurun_out
mkdir -p $OUT
for k in k_encode k_decode; do
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:^$k -c 1 -f -o $OUT/r2_final_$k \
        python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu --no-e2e --watchdog 350 > $OUT/r2_final_$k.log 2>&1
    ncu -i $OUT/r2_final_$k.ncu-rep --page raw --csv > $OUT/r2_final_${k}_raw.csv 2>/dev/null
done
for wl in c2 c3 c5; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum \
    --clock-control none -c 160 --csv --page raw --log-file $OUT/r2_final_allkernels_${wl}_raw.csv \
    python bench.py --workload $wl --steps 1 --warmup 0 --no-cpu --no-e2e --watchdog 500 > $OUT/r2_final_allkernels_$wl.log 2>&1
done
echo "ncu rc=$?"
