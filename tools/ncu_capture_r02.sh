#!/bin/bash
# Round-2 evidence capture on the GPU box (run under gpurun from the repo root):  bash tools/ncu_capture_r02.sh TAG
#   1. plain run of the profiled command (must exit 0 before anything runs under ncu)
#   2. `ncu --set full --import-source on` of every kernel of ONE full-size bf16 encoder layer, forward + backward, twice
#      (tools/layer_step.py: 1 114 112 rows = a 65 536-node timestep) -> per-kernel table / jsonl (tools/summarize_ncu.py)
#   3. SASS stall tables of the three FFN kernels
TAG=${1:-v1}
CMD="python tools/layer_step.py 65536 2"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_layer_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_layer_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_layer_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:"ffn_tc|ln_bwd|ln_fwd|attn_tc|gemm_tc" -c 40 -o gpurun_out/prof_layer_$TAG $CMD > gpurun_out/ncu_layer_$TAG.log 2>&1
ncu -i gpurun_out/prof_layer_$TAG.ncu-rep --page raw --csv > gpurun_out/r02_ncu_layer_${TAG}_raw.csv
python tools/summarize_ncu.py gpurun_out/r02_ncu_layer_${TAG}_raw.csv > gpurun_out/r02_ncu_layer_${TAG}_summary.jsonl 2> gpurun_out/r02_ncu_layer_${TAG}_table.txt
cat gpurun_out/r02_ncu_layer_${TAG}_table.txt
: > gpurun_out/r02_ncu_layer_${TAG}_stalls.txt
N=$(grep -c "ffn_tc\|ln_bwd\|ln_fwd\|attn_tc\|gemm_tc" gpurun_out/r02_ncu_layer_${TAG}_raw.csv)
for k in $(seq 0 $((N / 2 - 1))); do
  ncu -i gpurun_out/prof_layer_$TAG.ncu-rep --page source --csv --print-source sass --launch-skip $k --launch-count 1 > /tmp/sass_$k.csv 2>/dev/null
  NAME=$(grep -m1 'Kernel Name' /tmp/sass_$k.csv | cut -c1-120)
  case "$NAME" in *ffn_tc*) echo "== launch $k: $NAME" >> gpurun_out/r02_ncu_layer_${TAG}_stalls.txt; python tools/sass_stalls.py /tmp/sass_$k.csv 16 >> gpurun_out/r02_ncu_layer_${TAG}_stalls.txt 2>&1;; esac
done
SZ=$(du -sm gpurun_out | cut -f1); if [ "$SZ" -gt 50 ]; then rm -f gpurun_out/prof_layer_$TAG.ncu-rep; fi
du -sh gpurun_out
