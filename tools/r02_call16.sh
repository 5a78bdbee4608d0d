#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_numbers.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02l_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02l_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02l_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02l_smoke.txt
bash tools/gpu_ab.sh 10 "l_default:"
