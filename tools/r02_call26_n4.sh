#!/bin/bash
# Round 2, call 26 (4 GPUs, charged 4x): CTAs of the overlapped NVLS all-reduce at N = 4 (the rule 64 at N <= 2 / 32 above was
# extrapolated for N = 4).
#   /usr/local/graft/bin/gpurun --gpus 4 --timeout 400 -- 'bash tools/r02_call26_n4.sh'
set -u
N=4
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {  # label, flags
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $N --steps 30 --warmup 5 --no_cpu 1 $2 2> gpurun_out/c26_$1.err | grep '^{' > gpurun_out/c26_$1.json
  echo "$1 [$2] rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c26_$1.json')); print(d['value'], 'clips/s', d['ms_per_step'], 'ms/step', 'e2e ms', d['e2e']['ms_per_step'])" 2>&1 | tail -1)"
}
run c32 "--nvls_ctas 32"
run c64 "--nvls_ctas 64"
