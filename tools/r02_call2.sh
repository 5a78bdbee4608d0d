#!/bin/bash
# Round 2, call 2: GPU suites on the new build (device-side dropout stream, 4-wide hash, OCC=2 GEMM variants), then A/B of the
# two-CTAs-per-SM policy.   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r02_call2.sh'
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02b_pytest_gpu.txt
bash tools/gpu_ab.sh 10 \
  "b_default:" \
  "b_occ2_bert:--occ2 2 --occ2_gflop 12.9" \
  "b_occ2_14:--occ2 2 --occ2_gflop 14" \
  "b_occ2_30:--occ2 2 --occ2_gflop 30" \
  "b_occ2_45:--occ2 2 --occ2_gflop 45" \
  "b_occ2_all:--occ2 2" \
  "b_occ2_bert_pdl:--occ2 2 --occ2_gflop 12.9 --pdl 1" \
  "b_occ2_30_pdl:--occ2 2 --occ2_gflop 30 --pdl 1" \
  "b_occ2_all_1stream:--occ2 2 --overlap_wgrad 0" \
  "b_occ2_all_1stream_pdl:--occ2 2 --overlap_wgrad 0 --pdl 1"
