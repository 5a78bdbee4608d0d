#!/bin/bash
# N = 8 A/B of the exchange defaults (charged 8 x): gpurun --gpus 8 --timeout 900 -- 'bash tools/round2_multi_gpu8.sh'
set -u
N=8
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $N --steps 10 --warmup 3 --no_cpu 1 --optimizer 0 $2 2> gpurun_out/mg_${N}_$1.err | grep '^{' > gpurun_out/mg_${N}_$1.json
  echo "$1 [$2] $(python -c "import json; d=json.load(open('gpurun_out/mg_${N}_$1.json')); print(d['value'], 'clips/s', d['ms_per_step'], 'ms/step')" 2>&1 | tail -1)"
}
run base ""
run buckets "--cnn_buckets 1"
run nvls_buckets "--exchange nvls --cnn_buckets 1"
run bf16_buckets "--wire bf16 --cnn_buckets 1"
timeout 100 python bench.py --steps 10 --warmup 3 --no_cpu 1 --optimizer 0 2>/dev/null | grep '^{' > gpurun_out/mg_8_n1.json; python -c "import json; d=json.load(open('gpurun_out/mg_8_n1.json')); print('N=1 on this box', d['value'], d['ms_per_step'])"
