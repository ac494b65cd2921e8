#!/bin/bash
# every tile of the 1-D layers through the TMA ring (clip-end rows patched by the producers): parity + A/B
TAG=${1:-r2t}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_parity.py tests/test_gpu_fullshape.py -q -x -m gpu) > gpurun_out/pytest_${TAG}.txt 2>&1
tail -4 gpurun_out/pytest_${TAG}.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e"
for V in 1 0; do
  FCB_TC_TMA_ALL=$V timeout 120 $B --workload config2 > gpurun_out/bench_config2_all${V}_${TAG}.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_config2_all${V}_${TAG}.json"))
print("tma_all=${V}", round(d["ms_per_step"], 3), {k: round(v, 2) for k, v in d["phase_ms_last_step"].items()})
PY
done
