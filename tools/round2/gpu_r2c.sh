#!/bin/bash
# round-2 step: fp16-split tensor-core conv bring-up.  tools/gpu_r2c.sh <tag>
TAG=${1:-r2c}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_parity.py tests/test_gpu_freq.py -x -q) > gpurun_out/pytest_${TAG}.txt 2>&1
tail -15 gpurun_out/pytest_${TAG}.txt
for WL in config2 config4 config4_gr8; do
  timeout 300 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${WL}_${TAG}.json 2> gpurun_out/bench_${WL}_${TAG}.err
  tail -c 600 gpurun_out/bench_${WL}_${TAG}.json
done
FCB_LSTM_TRACE=1 timeout 300 python - > gpurun_out/lstm_trace_${TAG}.txt 2>&1 <<'PY'
import torch
from funcodec_b200 import get_config, init_state_dict
from funcodec_b200.encodec import B200Encodec
cfg = get_config("encodec_16k_n32_ds640")
m = B200Encodec(cfg, init_state_dict(cfg, 0), "cuda:0")
x = 0.1 * torch.randn(16, 160000, device="cuda")
for _ in range(3):
    m.inference(x, need_sub_quants=False)
torch.cuda.synchronize()
del m
PY
tail -70 gpurun_out/lstm_trace_${TAG}.txt | head -40
bash tools/gpu_knockout.sh ${TAG} config2 > gpurun_out/ko_${TAG}.log 2>&1
