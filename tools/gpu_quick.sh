#!/bin/bash
# quick regression + bench + launch list (no full ncu): tools/gpu_quick.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "phases", d["phase_ms_last_step"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None, "launches", d["gpu_launches"], d["clocks"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv > gpurun_out/launch_summary_${TAG}.txt 2>&1
head -70 gpurun_out/launch_summary_${TAG}.txt
