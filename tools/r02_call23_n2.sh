#!/bin/bash
# Round 2, call 23 (2 GPUs, charged 2x): the driver's own N = 2 launch line on HEAD's defaults (exchange auto -> NVLS, CNN buckets,
# gradient clearing inside the step) and the A/B of the new in-step gradient clearing under the exchange.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash tools/r02_call23_n2.sh'
set -u
N=2
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {  # label, flags
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $N --steps 20 --warmup 5 --no_cpu 1 $2 2> gpurun_out/c23_$1.err | grep '^{' > gpurun_out/c23_$1.json
  echo "$1 [$2] rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c23_$1.json')); print(d['value'], 'clips/s', d['ms_per_step'], 'ms/step', 'e2e', d['e2e']['value'], d['config'].get('exchange'))" 2>&1 | tail -1)"
}
run default ""
run zero_serial "--zero_grad_in_forward 0"
run default_again ""
timeout 120 python bench.py --steps 20 --warmup 5 --no_cpu 1 2>/dev/null | grep '^{' > gpurun_out/c23_n1.json; python -c "import json; d=json.load(open('gpurun_out/c23_n1.json')); print('N=1 on this box', d['value'], d['ms_per_step'])"
tail -3 gpurun_out/c23_default.err
