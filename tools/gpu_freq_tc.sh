#!/bin/bash
# 2-D tensor-core bring-up: full GPU suite (all failures listed), config-4 bench (TC vs SIMT), default bench.
TAG=${1:-r1p}
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_freq.py -m gpu -q -s -rf 2>&1 | grep -v "^$" | tail -120 > gpurun_out/pytest_freq_${TAG}.log
grep -c "use_tc2d" gpurun_out/pytest_freq_${TAG}.log; tail -25 gpurun_out/pytest_freq_${TAG}.log
timeout 240 python -m pytest tests -m gpu -q --deselect tests/test_gpu_freq.py 2>&1 | tail -5 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
timeout 150 python bench.py --workload config4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_${TAG}.json 2> gpurun_out/bench_cfg4_${TAG}.err
FCB_USE_TC2D=0 timeout 150 python bench.py --workload config4 --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/bench_cfg4_simt_${TAG}.json 2> gpurun_out/bench_cfg4_simt_${TAG}.err
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
for f in ("bench_cfg4_${TAG}", "bench_cfg4_simt_${TAG}", "bench_${TAG}"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value", d["value"], "ms/step", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "phases", d["phase_ms_last_step"], "roofline", d["roofline"]["frac"])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err" % f).read()[-1500:])
PY
