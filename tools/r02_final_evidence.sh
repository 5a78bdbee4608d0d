#!/bin/bash
# Evidence of the committed HEAD (one B200): GPU suites, smoke, bench lines of every BASELINE config, ncu launch list of the bench
# command, ncu --set full rows of the step's representative GEMM launches, per-kernel step breakdown.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/r02_final_evidence.sh'
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_numbers.txt gpurun_out/r02z_attn_full_summary.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02z_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02z_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02z_smoke.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02z_bench_headline.json 2> gpurun_out/r02z_bench_headline.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_bench_reference.json 2> gpurun_out/r02z_bench_reference.err; echo "reference rc=$?"
for c in c2 c3 c4 c5; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no_cpu 1 > gpurun_out/r02z_bench_$c.json 2> gpurun_out/r02z_bench_$c.err; echo "$c rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/r02z_launches.csv \
  python bench.py --steps 2 --warmup 1 --no_cpu 1 --optimizer 0 --graph 0 > gpurun_out/r02z_ncu_bench.log 2>&1; echo "launch list rc=$?"
python tools/summarize_ncu.py gpurun_out/r02z_launches.csv gpurun_out/r02z_launch_list_summary.txt > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kernel -c 8 -f -o gpurun_out/r02z_gemm_full \
  python tools/ncu_shapes.py --match "m2624 n3072 k768 mode0 t1 r0 a0 o1" "m2624 n768 k768 mode0 t1 r1" "m2624 n2304 k768 mode0" "m2624 n3072 k768 mode2 t1 r0 a1" \
  "m401408 n256 k64 mode0 t1 r1" "m32768 n256 k256 mode0 t9" "m10368 n2048 k768 mode2 t9" "m10368 n768 k2048 mode0 t9" > gpurun_out/r02z_gemm_full_stdout.txt 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r02z_gemm_full.ncu-rep --page raw --csv > gpurun_out/r02z_gemm_full_raw.csv 2>/dev/null
python tools/summarize_ncu_full.py gpurun_out/r02z_gemm_full_raw.csv > gpurun_out/r02z_gemm_full_summary.txt 2>&1
tail -9 gpurun_out/r02z_gemm_full_stdout.txt | head -8 >> gpurun_out/r02z_gemm_full_summary.txt
sz=$(stat -c %s gpurun_out/r02z_gemm_full.ncu-rep 2>/dev/null || echo 0); if [ "$sz" -gt 30000000 ]; then rm -f gpurun_out/r02z_gemm_full.ncu-rep; fi
# the attention kernels (mma.sync): the L = 41 forward / backward of the headline step, the flash forward of config 5 (L = 521)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_tc -c 3 -f -o gpurun_out/r02z_attn_full \
  python bench.py --steps 1 --warmup 1 --no_cpu 1 --optimizer 0 --graph 0 > gpurun_out/r02z_attn_full_stdout.txt 2>&1; echo "ncu attention rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd_flash -c 1 -f -o gpurun_out/r02z_attn_flash_full \
  python bench.py --config c5 --steps 1 --warmup 1 --no_cpu 1 --graph 0 > gpurun_out/r02z_attn_flash_full_stdout.txt 2>&1; echo "ncu flash rc=$?"
for f in r02z_attn_full r02z_attn_flash_full; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
  python tools/summarize_ncu_full.py gpurun_out/${f}_raw.csv >> gpurun_out/r02z_attn_full_summary.txt 2>&1
  sz=$(stat -c %s gpurun_out/$f.ncu-rep 2>/dev/null || echo 0); if [ "$sz" -gt 20000000 ]; then rm -f gpurun_out/$f.ncu-rep; fi
done
timeout 300 python tools/profile_step.py --out gpurun_out/r02z_step_breakdown.txt > /dev/null 2>&1; echo "breakdown rc=$?"
head -30 gpurun_out/r02z_launch_list_summary.txt
cat gpurun_out/r02z_gemm_full_summary.txt
cat gpurun_out/r02z_attn_full_summary.txt
