#!/bin/bash
# One kernel of the full-size bf16 / fp32 layer step with stall + executed-instruction tables:
#   bash tools/ncu_capture_one.sh KERNEL_REGEX SKIP TAG [bf16|fp32] [nodes]
K=$1; SKIP=${2:-0}; TAG=${3:-one}; PREC=${4:-bf16}; NODES=${5:-65536}
CMD="python tools/layer_step.py $NODES 1 $PREC"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip $SKIP --launch-count 1 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv --print-source sass > /tmp/sass_$TAG.csv 2>/dev/null
{ python tools/sass_exec_profile.py /tmp/sass_$TAG.csv 200; python tools/sass_stalls.py /tmp/sass_$TAG.csv 28; } > gpurun_out/r02_ncu_${TAG}_stalls.txt 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[2] if len(r)>2 else r[1]
for k in ['gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']:
    if k in h: print(k, v[h.index(k)])
" >> gpurun_out/r02_ncu_${TAG}_stalls.txt
rm -f gpurun_out/prof_$TAG.ncu-rep
cat gpurun_out/r02_ncu_${TAG}_stalls.txt
