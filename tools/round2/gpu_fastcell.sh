#!/bin/bash
# experiment: hardware ex2 / rcp LSTM gates (FCB_LSTM_FASTCELL=1) -- parity on the full shapes + config-2 / config-5 timing
TAG=${1:-r2fc}
mkdir -p gpurun_out
FCB_LSTM_FASTCELL=1 timeout 600 python -m pytest tests/test_gpu_fullshape.py tests/test_gpu_parity.py -q -m gpu > gpurun_out/pytest_fastcell_${TAG}.txt 2>&1
tail -4 gpurun_out/pytest_fastcell_${TAG}.txt
cp gpurun_out/parity_records.json gpurun_out/parity_records_fastcell_${TAG}.json 2>/dev/null
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e"
for V in 0 1; do
  FCB_LSTM_FASTCELL=$V timeout 120 $B --workload config2 > gpurun_out/bench_config2_fc${V}_${TAG}.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_config2_fc${V}_${TAG}.json"))
print("fastcell=${V}", round(d["ms_per_step"], 3), {k: round(v, 2) for k, v in d["phase_ms_last_step"].items()})
PY
done
