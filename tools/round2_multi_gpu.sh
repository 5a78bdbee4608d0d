#!/bin/bash
# N-GPU A/B of the data-parallel exchange switches (charged N x):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/round2_multi_gpu.sh 2'
set -u
N=${1:-2}
mkdir -p gpurun_out
run() {  # label, flags
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $N --steps 10 --warmup 3 --no_cpu 1 --optimizer 0 $2 > gpurun_out/mg_${N}_$1.json 2> gpurun_out/mg_${N}_$1.err
  echo "$1 [$2] rc=$? $(python -c "import json; d=json.load(open('gpurun_out/mg_${N}_$1.json')); print(d['value'], 'clips/s', d['ms_per_step'], 'ms/step')" 2>&1 | tail -1)"
}
run base ""
run buckets "--cnn_buckets 1"
run bf16 "--wire bf16"
run bf16_buckets "--wire bf16 --cnn_buckets 1"
run bf16_buckets_ctas16 "--wire bf16 --cnn_buckets 1 --nccl_ctas 16"
run nvls "--exchange nvls"
run nvls_buckets "--exchange nvls --cnn_buckets 1"
# stand-alone correctness + bandwidth of the NVLS kernel against NCCL
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) \
  tools/test_nvls.py > gpurun_out/mg_${N}_nvls_probe.json 2> gpurun_out/mg_${N}_nvls_probe.err; echo "nvls probe rc=$?"; tail -1 gpurun_out/mg_${N}_nvls_probe.json
timeout 120 python bench.py --steps 10 --warmup 3 --no_cpu 1 --optimizer 0 > gpurun_out/mg_1_base.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/mg_1_base.json')); print('N=1 on this box', d['value'], d['ms_per_step'])"
