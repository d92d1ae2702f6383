#!/bin/bash
# `ncu --set full --import-source on` of the FULL-SIZE FFN launches of one step (bench.py --nodes 65536: 1 114 112 rows):
# per step the FFN kernels launch as fwd F F F S | dgrad S wgrad S | (dgrad F wgrad F) x 3  (F = full timestep, S = last timestep).
TAG=${1:-v26}
CMD="python bench.py --steps 1 --warmup 3 --nodes 65536 --precision bf16 --no-cpu-baseline"
$CMD > gpurun_out/plain_ffn_$TAG.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"ffn_tc_(fwd|dgrad|wgrad)" -s 12 -c 1 -o gpurun_out/prof_ffnfwd_$TAG $CMD > gpurun_out/ncu_ffnfwd_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ffn_tc_(fwd|dgrad|wgrad)" -s 18 -c 2 -o gpurun_out/prof_ffnbwd_$TAG $CMD > gpurun_out/ncu_ffnbwd_$TAG.log 2>&1
for w in ffnfwd ffnbwd; do
  ncu -i gpurun_out/prof_${w}_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${w}_${TAG}_raw.csv
  python tools/summarize_ncu.py gpurun_out/ncu_${w}_${TAG}_raw.csv > gpurun_out/ncu_${w}_${TAG}_summary.jsonl 2> gpurun_out/ncu_${w}_${TAG}_table.txt
  cat gpurun_out/ncu_${w}_${TAG}_table.txt
done
for k in 0 1; do
  ncu -i gpurun_out/prof_ffnbwd_$TAG.ncu-rep --page source --csv --print-source sass --launch-skip $k --launch-count 1 > /tmp/sass_b$k.csv 2>/dev/null
  echo "== $(grep -m1 'Kernel Name' /tmp/sass_b$k.csv | cut -c1-120)" >> gpurun_out/ncu_ffn_${TAG}_stalls.txt
  python tools/sass_stalls.py /tmp/sass_b$k.csv 25 >> gpurun_out/ncu_ffn_${TAG}_stalls.txt 2>&1
done
ncu -i gpurun_out/prof_ffnfwd_$TAG.ncu-rep --page source --csv --print-source sass --launch-skip 0 --launch-count 1 > /tmp/sass_f.csv 2>/dev/null
echo "== $(grep -m1 'Kernel Name' /tmp/sass_f.csv | cut -c1-120)" >> gpurun_out/ncu_ffn_${TAG}_stalls.txt
python tools/sass_stalls.py /tmp/sass_f.csv 25 >> gpurun_out/ncu_ffn_${TAG}_stalls.txt 2>&1
du -sh gpurun_out
