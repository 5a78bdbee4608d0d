#!/bin/bash
# A/B bench runs on one B200. usage: tools/gpu_ab.sh <steps> "<label>:<bench flags>" ...   (writes gpurun_out/ab_<label>.json)
steps=${1:-10}; shift
mkdir -p gpurun_out
for spec in "$@"; do
  label=${spec%%:*}; flags=${spec#*:}
  timeout 300 python bench.py --steps $steps --warmup 3 --no_cpu 1 $flags > gpurun_out/ab_${label}.json 2> gpurun_out/ab_${label}.err
  echo "$label [$flags] rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/ab_${label}.json')); print(d['value'], 'clips/s', d['ms_per_step'], 'ms/step e2e', d['e2e']['value'], 'gemm_ms', d['roofline']['gemm_ms_per_step'], 'frac', d['roofline']['frac'], 'launches/step', d['gpu_launches_per_step'])" 2>&1 | tail -1)"
done
