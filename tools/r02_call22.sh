#!/bin/bash
# Round 2, call 22: non-GEMM tail (48-row attention kernels, position-major embedding backwards, 16-byte reductions, gradient
# clearing beside the transformer forward, PDL for the small kernels only): GPU suite, A/B bench lines, per-kernel step breakdown.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/r02_call22.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c22_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c22_pytest_gpu.txt
run() {  # name, flags...
  local name=$1; shift
  timeout 300 python bench.py --steps 20 --warmup 5 --no_cpu 1 "$@" 2> gpurun_out/c22_bench_$name.err | grep '^{' > gpurun_out/c22_bench_$name.json
  python - "$name" <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/c22_bench_%s.json" % n).read())
    print("%-28s %9.1f clips/s  %7.3f ms/step  e2e %9.1f  gemm frac %.4f  clocks %s" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print(n, "FAILED", e)
P
}
run default
run rows48_off --attn_rows48 0
run zero_serial --zero_grad_in_forward 0
run pdl2 --pdl 2
run old_paths --attn_rows48 0 --zero_grad_in_forward 0
run default_again
timeout 300 python tools/profile_step.py --out gpurun_out/c22_step_breakdown.txt > gpurun_out/c22_profile_step.log 2>&1; echo "breakdown rc=$?"; tail -3 gpurun_out/c22_profile_step.log
grep -v "^  gemm mode" gpurun_out/c22_step_breakdown.txt | head -60
