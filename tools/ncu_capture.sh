#!/bin/bash
# One `ncu --set full` capture of a few GEMM launches of a training step; exports the raw page as CSV and keeps the
# report only if it is small enough to travel back (gpurun_out/ is capped at 64 MiB).
# usage: tools/ncu_capture.sh <name> <skip> <count>
set -u
name=$1; skip=$2; count=$3
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kernel -s "$skip" -c "$count" \
    -f -o gpurun_out/$name python tools/profile_step.py --ncu 1 > gpurun_out/${name}_stdout.txt 2>&1
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
python tools/summarize_ncu_full.py gpurun_out/${name}_raw.csv > gpurun_out/${name}_summary.txt 2>&1
sz=$(stat -c %s gpurun_out/$name.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 20000000 ]; then rm -f gpurun_out/$name.ncu-rep; echo "report too large ($sz B): kept CSV only"; fi
tail -3 gpurun_out/${name}_stdout.txt
cat gpurun_out/${name}_summary.txt
