#!/bin/bash
# First hardware contact of the tcgen05 conv + persistent LSTM: layer tests with a hard timeout (a hung
# mbarrier pipeline must not eat the GPU budget), then the model-level suite.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_layers.py -q -x -k "simt" -s 2>&1 | tail -30 > gpurun_out/layers_simt.log
echo "--- simt rc=$?" >> gpurun_out/layers_simt.log
timeout 300 python -m pytest tests/test_gpu_layers.py -q -k "tc" -s 2>&1 | grep -E "^tc|passed|failed|Error|error|assert" | head -80 > gpurun_out/layers_tc.log
echo "--- tc rc=$?" >> gpurun_out/layers_tc.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -30 > gpurun_out/parity.log
tail -5 gpurun_out/layers_simt.log; cat gpurun_out/layers_tc.log; tail -15 gpurun_out/parity.log
