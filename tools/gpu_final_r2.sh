#!/bin/bash
# round-2 final: the full GPU suite, the default bench line, the other configs, ncu launch list + full metric set of one step's
# conv launches (-> profiles/conv_traffic.json) + RVQ/LSTM.   tools/gpu_final_r2.sh <tag>
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
(time timeout 900 python -m pytest tests -m gpu -q -rf) > gpurun_out/pytest_full_${TAG}.txt 2>&1
tail -6 gpurun_out/pytest_full_${TAG}.txt
cp gpurun_out/parity_records.json gpurun_out/parity_records_${TAG}.json 2>/dev/null
timeout 400 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 1500 gpurun_out/bench_${TAG}.json
for WL in config1 config3 config4 config4_gr8 config5; do
  timeout 200 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_${WL}_${TAG}.json 2> gpurun_out/bench_${WL}_${TAG}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${WL}_${TAG}.json"))
    print("${WL}", round(d["ms_per_step"], 2), "ms", round(d["value"]), "frames/s e2e", round(d["e2e"]["value"]), {k: round(v, 2) for k, v in d["phase_ms_last_step"].items()}, "frac", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("${WL} FAILED", e)
PY
done
BENCH="python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline --no-extras"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launch_${TAG}.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv > gpurun_out/launch_summary_${TAG}.txt 2>&1
head -20 gpurun_out/launch_summary_${TAG}.txt
NL=$(grep -c "conv1d" gpurun_out/launch_summary_${TAG}.txt)
# full metric set for the conv launches of the LAST step (3 warm-up steps + 1 timed: skip 3/4 of the conv launches)
NCONV=$(python - <<PY
import csv
rows = [l for l in open("gpurun_out/launches_${TAG}.csv") if not l.startswith("==")]
n = sum(1 for r in csv.DictReader(rows) if r.get("Metric Name") == "gpu__time_duration.sum" and "conv1d" in r["Kernel Name"])
print(n // 4)
PY
)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1d -s $((3 * NCONV)) -c $NCONV -f -o /tmp/conv_${TAG} $BENCH > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i /tmp/conv_${TAG}.ncu-rep --page raw --csv > /tmp/conv_raw_${TAG}.csv 2>/dev/null
python tools/summarize_ncu_raw.py /tmp/conv_raw_${TAG}.csv --traffic-json gpurun_out/conv_traffic.json config2 profiles/conv_ncu_${TAG}.txt > gpurun_out/conv_ncu_${TAG}.txt 2>&1
cat gpurun_out/conv_traffic.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"rvq_tc|lstm_seq" -s 15 -c 3 -f -o /tmp/rl_${TAG} $BENCH >> gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i /tmp/rl_${TAG}.ncu-rep --page raw --csv > /tmp/rl_raw_${TAG}.csv 2>/dev/null
python tools/summarize_ncu_raw.py /tmp/rl_raw_${TAG}.csv > gpurun_out/rvq_lstm_ncu_${TAG}.txt 2>&1
ls -la gpurun_out | tail -5
