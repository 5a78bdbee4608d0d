#!/bin/bash
# Round 2, call 24 (2 GPUs): CTAs of the overlapped NVLS all-reduce at N = 2 (each rank reduces 1/2 of the buffer - four times
# the per-rank work of N = 8, where 32 CTAs were best).
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash tools/r02_call24_n2.sh'
set -u
N=2
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {  # label, flags
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $N --steps 30 --warmup 5 --no_cpu 1 $2 2> gpurun_out/c24_$1.err | grep '^{' > gpurun_out/c24_$1.json
  echo "$1 [$2] rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c24_$1.json')); print(d['value'], 'clips/s', d['ms_per_step'], 'ms/step', 'e2e ms', d['e2e']['ms_per_step'])" 2>&1 | tail -1)"
}
run c32_a "--nvls_ctas 32"
run c64_a "--nvls_ctas 64"
run c96_a "--nvls_ctas 96"
run c32_b "--nvls_ctas 32"
run c64_b "--nvls_ctas 64"
run c96_b "--nvls_ctas 96"
run nccl "--exchange nccl"
