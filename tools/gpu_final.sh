#!/bin/bash
# round-end validation: full GPU suite (all failures listed), memcheck of the new paths, config-4 / default bench,
# config-4 launch list.  tools/gpu_final.sh <tag>
TAG=${1:-r1q}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -rf -s 2>&1 | grep -av "^$" > gpurun_out/pytest_full_${TAG}.log
grep -a "passed\|failed" gpurun_out/pytest_full_${TAG}.log | tail -3
grep -a "^FAILED\|^ERROR" gpurun_out/pytest_full_${TAG}.log | head -40
grep -a "segmented\|overlap-add\|encoder_out max\|decode-only" gpurun_out/pytest_full_${TAG}.log | head -20
timeout 120 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_freq.py tests/test_gpu_parity.py -q -k "(full_config and 7) or segmented_golden or (small_channels and 7 and model.13)" 2>&1 | tail -4 > gpurun_out/memcheck_${TAG}.log
cat gpurun_out/memcheck_${TAG}.log
timeout 100 python bench.py --workload config4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_${TAG}.json 2> gpurun_out/bench_cfg4_${TAG}.err
timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
for f in ("bench_cfg4_${TAG}", "bench_${TAG}"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value", d["value"], "ms/step", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "phases", d["phase_ms_last_step"], "roofline", d["roofline"]["frac"])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/%s.err" % f).read()[-1500:])
PY
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg4_${TAG}.csv python bench.py --workload config4 --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline > gpurun_out/ncu_launch_cfg4_${TAG}.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_cfg4_${TAG}.csv > gpurun_out/launch_summary_cfg4_${TAG}.txt 2>&1
head -40 gpurun_out/launch_summary_cfg4_${TAG}.txt
