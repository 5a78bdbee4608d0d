#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/r02g_pytest_ops.txt 2>&1; echo "pytest ops rc=$?"; tail -4 gpurun_out/r02g_pytest_ops.txt
timeout 300 python tools/probe_gemm_cta_timeline.py > gpurun_out/r02g_cta_timeline.txt 2>&1; echo "probe rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_ops.py > gpurun_out/r02g_pytest_rest.txt 2>&1; echo "pytest rest rc=$?"; tail -4 gpurun_out/r02g_pytest_rest.txt
bash tools/gpu_ab.sh 10 "g_default:" "g_occ2_14:--occ2 2 --occ2_gflop 14" "g_occ2_all:--occ2 2"
