#!/bin/bash
# multi-GPU validation: bench at N GPUs of one box incl. the config-5 extra block (NCCL scatter/gather + shared host batch).
N=${1:-2}; TAG=${2:-r2z}
mkdir -p gpurun_out
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_${TAG}.json 2> gpurun_out/bench_${N}gpu_${TAG}.err
echo "stdout lines: $(wc -l < gpurun_out/bench_${N}gpu_${TAG}.json)"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${N}gpu_${TAG}.json"))
print(d["n_gpus"], round(d["ms_per_step"], 2), round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["path"])
print(json.dumps(d.get("extra"), indent=1)[:2500])
PY
grep -c "NCCL INFO" gpurun_out/bench_${N}gpu_${TAG}.err; grep "NCCL INFO comm\|nranks" gpurun_out/bench_${N}gpu_${TAG}.err | head -4
tail -5 gpurun_out/bench_${N}gpu_${TAG}.err | cut -c1-300
