#!/bin/bash
# the weight_norm / causal branches first (fast, separate log), then the usual confirmation of the whole state.
TAG=${1:-r2n}
mkdir -p gpurun_out
(time timeout 240 python -m pytest tests -m gpu -q -rf -k "soundstream or weightnorm") > gpurun_out/pytest_norms_${TAG}.txt 2>&1
tail -15 gpurun_out/pytest_norms_${TAG}.txt
bash tools/gpu_confirm.sh ${TAG}
