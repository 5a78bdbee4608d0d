#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/r02h_pytest_ops.txt 2>&1; echo "pytest ops rc=$?"; tail -3 gpurun_out/r02h_pytest_ops.txt
timeout 300 python tools/probe_gemm_cta_timeline.py > gpurun_out/r02h_cta_timeline.txt 2>&1; echo "probe rc=$?"
bash tools/gpu_ab.sh 10 "h_default:" "h_pdl:--pdl 1" "h_1stream:--overlap_wgrad 0"
