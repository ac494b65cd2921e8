#!/bin/bash
# round-2 step e: TMA raw ring + 8-deep accumulator ring.  tools/gpu_r2e.sh <tag>
TAG=${1:-r2e}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_parity.py -x -q -m gpu) > gpurun_out/pytest_${TAG}.txt 2>&1
tail -8 gpurun_out/pytest_${TAG}.txt
if grep -q "failed\|error\|Error" gpurun_out/pytest_${TAG}.txt; then
  FCB_TC_TMA=0 timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_notma_${TAG}.txt 2>&1
  tail -8 gpurun_out/pytest_notma_${TAG}.txt
fi
timeout 300 python bench.py --workload config2 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_config2_${TAG}.json 2> gpurun_out/bench_config2_${TAG}.err
tail -c 300 gpurun_out/bench_config2_${TAG}.json
FCB_TC_TMA=0 timeout 300 python bench.py --workload config2 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e > gpurun_out/bench_config2_notma_${TAG}.json 2>/dev/null
tail -c 300 gpurun_out/bench_config2_notma_${TAG}.json
(time timeout 900 python -m pytest tests/test_gpu_freq.py tests/test_gpu_fullshape.py -x -q -m gpu) > gpurun_out/pytest2_${TAG}.txt 2>&1
tail -5 gpurun_out/pytest2_${TAG}.txt
for WL in config3 config5 config4_gr8; do
  timeout 300 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e > gpurun_out/bench_${WL}_${TAG}.json 2> gpurun_out/bench_${WL}_${TAG}.err
  tail -c 300 gpurun_out/bench_${WL}_${TAG}.json
done
bash tools/gpu_knockout.sh ${TAG} config2 > gpurun_out/ko_${TAG}.log 2>&1
