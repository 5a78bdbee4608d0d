#!/bin/bash
# A/B of the launch-level optimisations on one B200: programmatic dependent launch x clip batching.
# usage: tools/gpu_ab.sh [steps]   (writes gpurun_out/ab_*.json)
steps=${1:-10}
mkdir -p gpurun_out
for pdl in 0 1; do for cbat in 0 1; do
  timeout 300 python bench.py --steps $steps --warmup 3 --no_cpu 1 --pdl $pdl --clip_batching $cbat > gpurun_out/ab_pdl${pdl}_cb${cbat}.json 2> gpurun_out/ab_pdl${pdl}_cb${cbat}.err
  echo "pdl=$pdl clip_batching=$cbat rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/ab_pdl${pdl}_cb${cbat}.json')); print(d['value'], 'clips/s', d['ms_per_step'], 'ms/step e2e', d['e2e']['value'], 'gemm_ms', d['roofline']['gemm_ms_per_step'], 'frac', d['roofline']['frac'], 'launches/step', d['gpu_launches_per_step'])" 2>&1 | tail -1)"
done; done
