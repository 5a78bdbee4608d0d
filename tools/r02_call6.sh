#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02f_pytest_gpu.txt
timeout 300 python tools/probe_gemm_cta_timeline.py > gpurun_out/r02f_cta_timeline.txt 2>&1; echo "probe rc=$?"
bash tools/gpu_ab.sh 10 "f_default:" "f_occ2_14:--occ2 2 --occ2_gflop 14" "f_pdl:--pdl 1"
