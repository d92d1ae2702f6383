#!/bin/bash
# Round-1 evidence capture on the GPU box (run under gpurun from the repo root):  bash tools/ncu_capture.sh TAG
#   1. plain bench line of the profiled command (must exit 0 before anything runs under ncu)
#   2. launch list (gpu__time_duration only) of one 65 536-node run
#   3. `ncu --set full` of ONE whole train step (60 launches), summarised per kernel by tools/summarize_ncu.py
#   4. `ncu --set full --import-source on` of the FFN kernels (the roofline.traffic source) + SASS stall tables
TAG=${1:-v21}
CMD="python bench.py --steps 1 --warmup 3 --nodes 65536 --precision bf16 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > /dev/null 2>&1
python tools/agg_launches.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_${TAG}_summary.txt 2>/dev/null
ncu --set full --clock-control none -s 153 -c 60 -o gpurun_out/prof_step_$TAG $CMD > gpurun_out/ncu_step_$TAG.log 2>&1
ncu -i gpurun_out/prof_step_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_step_${TAG}_raw.csv
python tools/summarize_ncu.py gpurun_out/ncu_step_${TAG}_raw.csv > gpurun_out/ncu_step_${TAG}_summary.jsonl 2> gpurun_out/ncu_step_${TAG}_table.txt
cat gpurun_out/ncu_step_${TAG}_table.txt
rm -f gpurun_out/prof_step_$TAG.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"ffn_tc_(fwd|dgrad|wgrad)|gemm_tc_dgrad_wgrad|gemm_tc_rows" -s 36 -c 12 -o gpurun_out/prof_top_$TAG $CMD > gpurun_out/ncu_top_$TAG.log 2>&1
ncu -i gpurun_out/prof_top_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_top_${TAG}_raw.csv
python tools/summarize_ncu.py gpurun_out/ncu_top_${TAG}_raw.csv > gpurun_out/ncu_top_${TAG}_summary.jsonl 2> gpurun_out/ncu_top_${TAG}_table.txt
cat gpurun_out/ncu_top_${TAG}_table.txt
for k in 0 1 2 3 4 5 6 7 8 9 10 11; do
  ncu -i gpurun_out/prof_top_$TAG.ncu-rep --page source --csv --print-source sass --launch-skip $k --launch-count 1 > /tmp/sass_$k.csv 2>/dev/null
  echo "== launch $k: $(grep -m1 'Kernel Name' /tmp/sass_$k.csv | cut -c1-120)" >> gpurun_out/ncu_top_${TAG}_stalls.txt
  python tools/sass_stalls.py /tmp/sass_$k.csv 12 >> gpurun_out/ncu_top_${TAG}_stalls.txt 2>&1
done
SZ=$(du -sm gpurun_out | cut -f1); if [ "$SZ" -gt 55 ]; then rm -f gpurun_out/prof_top_$TAG.ncu-rep; fi
du -sh gpurun_out
