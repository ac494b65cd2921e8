#!/bin/bash
# Hardware bring-up of the tcgen05 conv: layer tests under both descriptor variants with hard timeouts
# (a hung mbarrier pipeline must not eat the GPU budget), then the model-level suite.
mkdir -p gpurun_out
for mode in 0 1; do
  FCB_TC_DBG_MODE=$mode timeout 300 python -m pytest tests/test_gpu_layers.py -q -k "tc" -s 2>&1 | grep -E "tc  |passed|failed|rror" | grep -v "^E  \|^>" | head -60 > gpurun_out/layers_tc_mode$mode.log
  echo "--- mode $mode rc=$?" >> gpurun_out/layers_tc_mode$mode.log
done
cat gpurun_out/layers_tc_mode0.log gpurun_out/layers_tc_mode1.log
