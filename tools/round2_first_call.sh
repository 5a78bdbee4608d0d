#!/bin/bash
# First gpurun call of the next round (one B200, ~12 min): everything that was written after round 1's GPU budget ran out.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# 1. the GPU suites (the late-sorted file holds the tests that have never run on a GPU), 2. the default bench line,
# 3. A/B of the switches that are off until measured, 4. a fresh launch list. Results land in gpurun_out/.
set -u
unset CB_EXPERIMENTAL          # 1st pass: the suite as the round-end driver runs it; 2nd pass adds the off-by-default kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.txt
CB_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zz_gpu_round1c.py -m gpu -q -s > gpurun_out/r02_pytest_zz.txt 2>&1; echo "zz rc=$?"; tail -5 gpurun_out/r02_pytest_zz.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"
bash tools/gpu_ab.sh 10 \
  "recast:--recast_in_step 1" \
  "default:" \
  "fusedloss:--fused_loss 1" \
  "fusedloss_pdl:--fused_loss 1 --pdl 1" \
  "pdl_late:--fused_loss 1 --pdl 1 --pdl_late 1" \
  "pdl_late_1stream:--fused_loss 1 --pdl 1 --pdl_late 1 --overlap_wgrad 0" \
  "mn3d:--mn3d 1" \
  "shortcut:--overlap_shortcut 1" \
  "sm144:--sm_limit 144" \
  "direct:--direct_store 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --no_cpu 1 --optimizer 0 > gpurun_out/r02_ncu_bench.log 2>&1
for f in 0 1; do timeout 300 python tools/bench_inference.py --flash $f > gpurun_out/r02_infer_flash$f.json 2> gpurun_out/r02_infer_flash$f.err; tail -1 gpurun_out/r02_infer_flash$f.json; done
python tools/summarize_ncu.py gpurun_out/r02_launches.csv > gpurun_out/r02_launch_list_summary.txt 2>&1; head -30 gpurun_out/r02_launch_list_summary.txt
