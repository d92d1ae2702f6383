CMD="python bench.py --steps 2 --warmup 3 --nodes 65536 --precision bf16 --no-cpu-baseline"
$CMD > gpurun_out/plain_v14.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"ffn_tc_(fwd|dgrad|wgrad)" -s 30 -c 7 -o gpurun_out/prof_ffn_v14 $CMD > gpurun_out/ncu_ffn_v14.log 2>&1
ls -la gpurun_out/prof_ffn_v14.ncu-rep
ncu -i gpurun_out/prof_ffn_v14.ncu-rep --page raw --csv > gpurun_out/ncu_ffn_v14_raw.csv
python tools/summarize_ncu.py gpurun_out/ncu_ffn_v14_raw.csv > gpurun_out/ncu_ffn_v14_summary.jsonl 2> gpurun_out/ncu_ffn_v14_table.txt
cat gpurun_out/ncu_ffn_v14_table.txt
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none -s 340 -c 90 -o gpurun_out/prof_step_v14 $CMD > gpurun_out/ncu_step_v14.log 2>&1
ls -la gpurun_out/prof_step_v14.ncu-rep
ncu -i gpurun_out/prof_step_v14.ncu-rep --page raw --csv > gpurun_out/ncu_step_v14_raw.csv
python tools/summarize_ncu.py gpurun_out/ncu_step_v14_raw.csv > gpurun_out/ncu_step_v14_summary.jsonl 2> gpurun_out/ncu_step_v14_table.txt
cat gpurun_out/ncu_step_v14_table.txt
# keep what travels back under the 64 MiB limit
SZ=$(du -sm gpurun_out | cut -f1); if [ "$SZ" -gt 55 ]; then rm -f gpurun_out/prof_step_v14.ncu-rep; fi
SZ=$(du -sm gpurun_out | cut -f1); if [ "$SZ" -gt 55 ]; then rm -f gpurun_out/prof_ffn_v14.ncu-rep; fi
du -sh gpurun_out
