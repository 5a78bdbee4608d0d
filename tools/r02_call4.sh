#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02d_pytest_gpu.txt
timeout 300 python tools/probe_gemm_cta_timeline.py > gpurun_out/r02d_cta_timeline.txt 2>&1; echo "probe rc=$?"
bash tools/gpu_ab.sh 10 \
  "d_default:" \
  "d_occ2_bert:--occ2 2 --occ2_gflop 12.9" \
  "d_occ2_14:--occ2 2 --occ2_gflop 14" \
  "d_occ2_all:--occ2 2" \
  "d_pdl:--pdl 1" \
  "d_1stream:--overlap_wgrad 0"
