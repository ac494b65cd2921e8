#!/bin/bash
TAG=${1:-r2m}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_parity.py tests/test_gpu_freq.py tests/test_gpu_fullshape.py -x -q -m gpu) > gpurun_out/pytest_${TAG}.txt 2>&1
tail -5 gpurun_out/pytest_${TAG}.txt
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e"
run() { name=$1; shift; env "$@" timeout 120 $B --workload $WL > gpurun_out/bench_${WL}_${name}_${TAG}.json 2>/dev/null; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${WL}_${name}_${TAG}.json"))
    print("${WL} ${name}", round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["phase_ms_last_step"].items()}, d["gpu_launches"])
except Exception as e:
    print("${WL} ${name} FAILED", e)
PY
}
WL=config2; run default A=1; run nofuse FCB_FUSE_STATS=0
WL=config1; run default A=1; run nofuse FCB_FUSE_STATS=0
WL=config4_gr8; run default A=1
