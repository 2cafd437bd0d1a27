#!/bin/bash
# usage: ncu_table.sh <out.csv> [bench args]
out=$1; shift
python bench.py --steps 2 --warmup 3 --no-cpu --graph 0 --tiles 1 "$@" > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 90 -c 60 --csv --log-file gpurun_out/$out python bench.py --steps 2 --warmup 3 --no-cpu --graph 0 --tiles 1 "$@" > gpurun_out/ncu.log 2>&1
echo done
