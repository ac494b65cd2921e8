#!/bin/bash
# minimal check of the deeper weight ring (config-4 paths only)
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gpu_freq.py -q -x 2>&1 | tail -3 > gpurun_out/pytest_mini.log
cat gpurun_out/pytest_mini.log
timeout 60 python bench.py --workload config4 --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/bench_cfg4_mini.json 2> gpurun_out/bench_cfg4_mini.err
python -c "
import json
d = json.load(open('gpurun_out/bench_cfg4_mini.json')); print('cfg4 deep ring', d['ms_per_step'], d['phase_ms_last_step'])"
