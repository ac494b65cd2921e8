#!/bin/bash
# last confirmation of the committed state: smoke, the whole GPU suite, the default bench line.
TAG=${1:-r2zz}
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
(time timeout 900 python -m pytest tests -m gpu -q -rf) > gpurun_out/pytest_full_${TAG}.txt 2>&1
tail -5 gpurun_out/pytest_full_${TAG}.txt
cp gpurun_out/parity_records.json gpurun_out/parity_records_${TAG}.json 2>/dev/null
timeout 400 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json"))
print(round(d["ms_per_step"], 3), round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"], d["extra"])
PY
