#!/bin/bash
# source-level stall profile of single launches: tools/gpu_src_profile.sh <tag> <kernel-regex> <skip> [<skip2> ...]
TAG=$1; KRE=$2; shift 2
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline"
for SK in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SK -c 1 -f -o /tmp/src_${TAG}_${SK} $BENCH > gpurun_out/ncu_src_${TAG}_${SK}.log 2>&1
  ncu -i /tmp/src_${TAG}_${SK}.ncu-rep --page raw --csv > /tmp/src_raw.csv 2>/dev/null
  python tools/summarize_ncu_raw.py /tmp/src_raw.csv | cut -c1-1200 > gpurun_out/src_${TAG}_${SK}.txt 2>&1
  ncu -i /tmp/src_${TAG}_${SK}.ncu-rep --page source --csv > /tmp/src_src.csv 2>/dev/null
  python tools/summarize_ncu_source.py /tmp/src_src.csv | head -60 >> gpurun_out/src_${TAG}_${SK}.txt 2>&1
done
