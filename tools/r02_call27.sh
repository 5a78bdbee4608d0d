#!/bin/bash
# Round 2, call 27: long-sequence attention forward with 32 query rows per warp (variant 2) against the default (variant 1).
#   /usr/local/graft/bin/gpurun --timeout 400 -- 'bash tools/r02_call27.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests -m gpu -q -x -k "long_sequences or config5" > gpurun_out/c27_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c27_pytest.txt
run() {
  local name=$1; shift
  timeout 120 python bench.py --config c5 --steps 20 --warmup 5 --no_cpu 1 "$@" 2> gpurun_out/c27_$name.err | grep '^{' > gpurun_out/c27_$name.json
  python -c "import json; d=json.load(open('gpurun_out/c27_$name.json')); print('$name', d['value'], 'clips/s', d['ms_per_step'], 'ms/step')" 2>&1 | tail -1
}
run v2 --attn_flash_pipe 2
run v1 --attn_flash_pipe 1
run v2_again --attn_flash_pipe 2
