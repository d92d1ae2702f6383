#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the tcgen05 / mbarrier kernels on small shapes.
# NOTE: on this GPU pool compute-sanitizer is closed by the operators (profiles/r02_sanitizer_unavailable.txt); the bounds checks
# the repo relies on instead are the guard-band tests (tests/test_gpu_tc.py::test_ffn_kernels_write_nothing_outside_their_outputs,
# tests/test_gpu_parity.py::test_sampled_softmax_label_out_of_range_is_reported_not_read).  On a box that allows it:
# Run on the GPU box:  bash tools/sanitize.sh   -> gpurun_out/r02_sanitizer_{memcheck,racecheck}.txt
set -u
mkdir -p gpurun_out
SEL='ffn_tc_forward and (256-64-256 or 37-64-128 or 300-7-256) or ffn_tc_backward and (256-64-128 or 37-64-256 or 300-7-256) or forward_mask and 129 or tile_images and 5-0.5 or tcgen05_operand_paths or gemm_tc_rows and 300 or inproj_attention and 77 or dgrad_wgrad_equals and 5000'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_tc.py -x -q -k "$SEL" > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r02_sanitizer_$tool.log
  { echo "== compute-sanitizer --tool $tool, pytest tests/test_gpu_tc.py -k \"$SEL\""; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit |Error|Hazard" gpurun_out/r02_sanitizer_$tool.log | head -40; } > gpurun_out/r02_sanitizer_$tool.txt
  cat gpurun_out/r02_sanitizer_$tool.txt
done
