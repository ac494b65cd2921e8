#!/bin/bash
# round-2 step h: LSTM compute-warp sets.  tools/gpu_r2h.sh <tag>
TAG=${1:-r2h}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_parity.py tests/test_gpu_freq.py -x -q -m gpu) > gpurun_out/pytest_${TAG}.txt 2>&1
tail -4 gpurun_out/pytest_${TAG}.txt
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e"
run() { name=$1; shift; env "$@" timeout 120 $B --workload $WL > gpurun_out/bench_${WL}_${name}_${TAG}.json 2>/dev/null; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${WL}_${name}_${TAG}.json"))
    print("${WL} ${name}", round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["phase_ms_last_step"].items()})
except Exception as e:
    print("${WL} ${name} FAILED", e)
PY
}
WL=config2; run default A=1; run nset2 FCB_LSTM_NSET=2
WL=config4_gr8; run default A=1; run nset1 FCB_LSTM_NSET=1; run nset2p3 FCB_LSTM_PAIRS=3
WL=config3; run default A=1; run nset1 FCB_LSTM_NSET=1
WL=config5; run default A=1
FCB_LSTM_TRACE=1 timeout 120 python tools/lstm_trace.py encodec_16k_n32_ds320 64 64000 > gpurun_out/lstm_trace_cfg3_${TAG}.txt 2>&1
