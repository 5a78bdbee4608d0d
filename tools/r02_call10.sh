#!/bin/bash
# round 2, call 10: full GPU suite with the round-2 parity tests, autotune of every BASELINE config (merged table), bench lines
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_numbers.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02j_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02j_pytest_gpu.txt
cp clipbert_b200/gemm_tuning.json gpurun_out/gemm_tuning.json
for c in headline c2 c3 c4 c5; do
  timeout 600 python tools/autotune_gemm.py --config $c --out gpurun_out/gemm_tuning.json --merge 1 > gpurun_out/r02j_autotune_$c.log 2>&1; echo "autotune $c rc=$? $(tail -1 gpurun_out/r02j_autotune_$c.log)"
done
cp gpurun_out/gemm_tuning.json clipbert_b200/gemm_tuning.json
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench_headline.json 2> gpurun_out/r02j_bench_headline.err; echo "bench rc=$?"
for c in c2 c3 c4 c5; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no_cpu 1 > gpurun_out/r02j_bench_$c.json 2> gpurun_out/r02j_bench_$c.err
  echo "$c rc=$? $(python -c "import json; d=json.load(open('gpurun_out/r02j_bench_$c.json')); print(d['value'], 'clips/s', d['ms_per_step'], 'ms/step frac', d['roofline'].get('frac'), 'whole', d['roofline'].get('whole_step_frac'))" 2>&1 | tail -1)"
done
python -c "import json; d=json.load(open('gpurun_out/r02j_bench_headline.json')); print('headline', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['gemm_ms_per_step'], d['roofline']['frac'])"
