#!/bin/bash
# fp32-mode evidence (run under gpurun from the repo root):  bash tools/ncu_capture_split.sh TAG
#   `ncu --set full --import-source on` of the bf16-split tcgen05 GEMMs of ONE full-size fp32 encoder layer (tools/layer_step.py ... fp32):
#   rows launches in order = in_proj, out_proj, linear1, linear2, dPre, dy1, dctx, dx; then the weight-gradient launches
TAG=${1:-v1}
NODES=${2:-32768}
CMD="python tools/layer_step.py $NODES 1 fp32"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_split_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_split_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_split_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:"gemm_split_rows" -c 8 -o gpurun_out/prof_split_$TAG $CMD > gpurun_out/ncu_split_$TAG.log 2>&1
ncu -i gpurun_out/prof_split_$TAG.ncu-rep --page raw --csv > gpurun_out/r02_ncu_split_${TAG}_raw.csv
python tools/summarize_ncu.py gpurun_out/r02_ncu_split_${TAG}_raw.csv > gpurun_out/r02_ncu_split_${TAG}_summary.jsonl 2> gpurun_out/r02_ncu_split_${TAG}_table.txt
cat gpurun_out/r02_ncu_split_${TAG}_table.txt
: > gpurun_out/r02_ncu_split_${TAG}_stalls.txt
for k in 2 3 4 5; do
  ncu -i gpurun_out/prof_split_$TAG.ncu-rep --page source --csv --print-source sass --launch-skip $k --launch-count 1 > /tmp/sass_$k.csv 2>/dev/null
  echo "== rows launch $k (0 in_proj, 1 out_proj, 2 linear1, 3 linear2, 4 dPre, 5 dy1)" >> gpurun_out/r02_ncu_split_${TAG}_stalls.txt
  python tools/sass_stalls.py /tmp/sass_$k.csv 24 >> gpurun_out/r02_ncu_split_${TAG}_stalls.txt 2>&1
done
rm -f gpurun_out/prof_split_$TAG.ncu-rep
du -sh gpurun_out
