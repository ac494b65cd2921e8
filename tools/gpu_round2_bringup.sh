#!/bin/bash
# First GPU call of round 2: validate and time the code paths written ahead of hardware access (all OFF by default):
#   tc_stage (cp.async-staged conv producers), conv2d_small_cout (halo-tile 32->3 conv), stft_tc (STFT / iSTFT as GEMMs),
#   lstm_prefetch_poll (software-pipelined barrier polling in the LSTM loader warp), tc_m256 (deep conv layers, M = 256).
# tools/gpu_round2_bringup.sh <tag>      (~2-3 min of box time)
TAG=${1:-r2a}
mkdir -p gpurun_out
FCB_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_gpu_experimental.py -q -s -rf 2>&1 | grep -av "^$" | tail -80 > gpurun_out/pytest_experimental_${TAG}.log
grep -a "passed\|failed\|^FAILED\|rel err\|max-abs err" gpurun_out/pytest_experimental_${TAG}.log | tail -60
run() {  # name, env..., bench args
    local name=$1; shift
    env "$@" timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e ${BENCH_ARGS} > gpurun_out/bench_${name}_${TAG}.json 2> gpurun_out/bench_${name}_${TAG}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${name}_${TAG}.json"))
    print("${name}", "ms/step", round(d["ms_per_step"], 3), "phases", {k: round(v, 2) for k, v in d["phase_ms_last_step"].items()})
except Exception as e:
    print("${name}", "FAILED", e); print(open("gpurun_out/bench_${name}_${TAG}.err").read()[-800:])
PY
}
BENCH_ARGS="" run cfg2_default FCB_TC_STAGE=0
BENCH_ARGS="" run cfg2_stage FCB_TC_STAGE=1
BENCH_ARGS="" run cfg2_stage_d2na4 FCB_TC_STAGE=1 FCB_TC_STAGE_DEPTH=2 FCB_TC_STAGE_NA=4
BENCH_ARGS="" run cfg2_stage_d3na2 FCB_TC_STAGE=1 FCB_TC_STAGE_DEPTH=3 FCB_TC_STAGE_NA=2
BENCH_ARGS="" run cfg2_stage_d4na2 FCB_TC_STAGE=1 FCB_TC_STAGE_DEPTH=4 FCB_TC_STAGE_NA=2
BENCH_ARGS="" run cfg2_m256 FCB_TC_M256=1
BENCH_ARGS="" run cfg2_lstm_prefetch FCB_LSTM_PREFETCH_POLL=1
BENCH_ARGS="" run cfg2_all FCB_TC_STAGE=1 FCB_TC_M256=1 FCB_LSTM_PREFETCH_POLL=1
BENCH_ARGS="--workload config4" run cfg4_default FCB_TC_STAGE=0
BENCH_ARGS="--workload config4" run cfg4_lstm_prefetch FCB_LSTM_PREFETCH_POLL=1
BENCH_ARGS="--workload config4" run cfg4_stage FCB_TC_STAGE=1
BENCH_ARGS="--workload config4" run cfg4_smallcout FCB_CONV2D_SMALL_COUT=1
BENCH_ARGS="--workload config4" run cfg4_stft FCB_STFT_TC=1
BENCH_ARGS="--workload config4" run cfg4_all FCB_TC_STAGE=1 FCB_CONV2D_SMALL_COUT=1 FCB_STFT_TC=1
FCB_TC_STAGE=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_stage_${TAG}.csv python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline > gpurun_out/ncu_launch_stage_${TAG}.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_stage_${TAG}.csv > gpurun_out/launch_summary_stage_${TAG}.txt 2>&1
head -70 gpurun_out/launch_summary_stage_${TAG}.txt
