#!/bin/bash
# Round captures of the cfg2 step (4096 random 8-node paths), run under gpurun from the repository root:
#   bash profiles/tools/capture.sh r02        -> gpurun_out/r02_ncu_launches.csv   (launch list: time, DRAM bytes, instructions ...)
#                                                gpurun_out/r02_full.ncu-rep        (--set full of the velocity / time kernels)
# Afterwards, on the CPU box:  python profiles/tools/ktable_json.py gpurun_out/r02_ncu_launches.csv profiles/r02_kernel_table.json 6544.7 <commit>
tag=${1:-r02}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
CMD="python bench.py --profile-step --steps 2 --warmup 2"
$CMD > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --metrics $M --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/${tag}_ncu_launches.csv $CMD > gpurun_out/${tag}_ncu.log 2>&1
$CMD > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_sample_prepass|k_fwd_chunked|k_bwd_chunked|k_time_state|k_build_props|k_time_sample" -s 12 -c 6 -f -o gpurun_out/${tag}_full $CMD >> gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
