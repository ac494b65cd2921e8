#!/bin/bash
# Where does the tensor-core conv kernel spend its time?  Knock out one pipeline stage at a time (FCB_TC_DBG mask, results are
# wrong on purpose) and compare per-launch durations.   tools/gpu_knockout.sh <tag> [workload]
TAG=${1:-ko}; WL=${2:-config2}
mkdir -p gpurun_out
MASKS="0 7 8 32 63 1023"
for M in $MASKS; do
  FCB_TC_DBG=$M timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ko${M}_${TAG}.csv \
    python bench.py --workload $WL --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline > gpurun_out/ncu_ko${M}_${TAG}.log 2>&1
done
python tools/summarize_knockouts.py $TAG $MASKS > gpurun_out/knockouts_${TAG}.txt 2>&1
cat gpurun_out/knockouts_${TAG}.txt
