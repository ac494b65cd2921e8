#!/bin/bash
# Profiling pass run on the GPU box through gpurun.  Writes SMALL artefacts into gpurun_out/ (the
# .ncu-rep files stay in /tmp on the box: gpurun only copies back <= 64 MiB).
#   usage: tools/gpu_profile.sh <tag> [conv_skip conv_count]
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline"
# 1. every launch with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launch_${TAG}.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv > gpurun_out/launch_summary_${TAG}.txt 2>&1
# 2. full metric set for one step's conv launches; only the raw CSV travels back
SKIP=${2:-144}; CNT=${3:-48}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv1d -s $SKIP -c $CNT -f -o /tmp/conv_${TAG} $BENCH > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i /tmp/conv_${TAG}.ncu-rep --page raw --csv > /tmp/conv_raw_${TAG}.csv 2>/dev/null
python tools/summarize_ncu_raw.py /tmp/conv_raw_${TAG}.csv > gpurun_out/conv_ncu_${TAG}.txt 2>&1
# 3. RVQ + one LSTM step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rvq_tc|lstm_seq" -s ${4:-15} -c 3 -f -o /tmp/rl_${TAG} $BENCH >> gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i /tmp/rl_${TAG}.ncu-rep --page raw --csv > /tmp/rl_raw_${TAG}.csv 2>/dev/null
python tools/summarize_ncu_raw.py /tmp/rl_raw_${TAG}.csv > gpurun_out/rvq_lstm_ncu_${TAG}.txt 2>&1
ncu -i /tmp/rl_${TAG}.ncu-rep --page source --csv > /tmp/rl_src_${TAG}.csv 2>/dev/null
python tools/summarize_ncu_source.py /tmp/rl_src_${TAG}.csv > gpurun_out/rvq_lstm_src_${TAG}.txt 2>&1
ls -la gpurun_out
