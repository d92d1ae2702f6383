CMD="python tools/layer_step.py 32768 1 fp32"
ncu --set full --clock-control none --import-source on -k regex:"gemm_split_rows" --launch-skip 2 --launch-count 1 -o gpurun_out/prof_l1 $CMD > gpurun_out/ncu_l1.log 2>&1
ncu -i gpurun_out/prof_l1.ncu-rep --page source --csv --print-source sass > /tmp/sass_l1.csv 2>/dev/null
python tools/sass_exec_profile.py /tmp/sass_l1.csv 150
python tools/sass_stalls.py /tmp/sass_l1.csv 14
ncu -i gpurun_out/prof_l1.ncu-rep --page raw --csv | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[2] if len(r)>2 else r[1]
for k in ['gpu__time_duration.sum','sm__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','launch__registers_per_thread','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed_op_shared_st.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']:
    if k in h: print(k, v[h.index(k)])
"
rm -f gpurun_out/prof_l1.ncu-rep
