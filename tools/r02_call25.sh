#!/bin/bash
# Round 2, call 25: the cp.async double-buffered long-sequence attention forward: its GPU tests, config 5 with / without it.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/r02_call25.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -m gpu -q -x -k "attention or long_sequences or config5 or native_resolution or long_text" > gpurun_out/c25_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c25_pytest.txt
run() {
  local name=$1; shift
  timeout 300 python bench.py --config c5 --steps 20 --warmup 5 --no_cpu 1 "$@" 2> gpurun_out/c25_$name.err | grep '^{' > gpurun_out/c25_$name.json
  python -c "import json; d=json.load(open('gpurun_out/c25_$name.json')); print('$name', d['value'], 'clips/s', d['ms_per_step'], 'ms/step', 'gemm ms', d['roofline']['gemm_ms_per_step'])" 2>&1 | tail -1
}
run pipe
run sync --attn_flash_pipe 0
run pipe_again
run sync_again --attn_flash_pipe 0
