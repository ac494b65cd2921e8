#!/bin/bash
# round-2 step d: coalesced epilogue stores, LSTM 4-clip groups / ring per group, skeleton knock-outs.  tools/gpu_r2d.sh <tag>
TAG=${1:-r2d}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_layers.py tests/test_gpu_parity.py tests/test_gpu_freq.py tests/test_gpu_fullshape.py tests/test_cli.py -x -q -m gpu) > gpurun_out/pytest_${TAG}.txt 2>&1
tail -15 gpurun_out/pytest_${TAG}.txt
cp gpurun_out/parity_records.json gpurun_out/parity_records_${TAG}.json 2>/dev/null
for WL in config2 config3 config4_gr8 config5; do
  timeout 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_${WL}_${TAG}.json 2> gpurun_out/bench_${WL}_${TAG}.err
  tail -c 400 gpurun_out/bench_${WL}_${TAG}.json
done
FCB_LSTM_GB=8 timeout 300 python bench.py --workload config2 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e > gpurun_out/bench_config2_gb8_${TAG}.json 2>/dev/null
FCB_LSTM_NBUF=2 timeout 300 python bench.py --workload config3 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e > gpurun_out/bench_config3_nbuf2_${TAG}.json 2>/dev/null
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
bash tools/gpu_knockout.sh ${TAG} config2 > gpurun_out/ko_${TAG}.log 2>&1
