#!/bin/bash
# all-kernel launch list of one config-2 step: duration, DRAM bytes, instructions
OUT=gpurun_out; TAG=${1:-r2_kl}
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum \
    --clock-control none -c 140 --csv --page raw --log-file $OUT/${TAG}_allkernels_raw.csv \
    python bench.py --workload ${2:-c2} --steps 1 --warmup 0 --no-cpu --no-e2e --watchdog 500 > $OUT/${TAG}_allkernels.log 2>&1
echo "ncu rc=$?"
