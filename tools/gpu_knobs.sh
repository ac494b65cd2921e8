#!/bin/bash
# A/B the tensor-core conv tuning knobs with the default bench (no cpu baseline, no e2e): prints ms/step + conv phases
mkdir -p gpurun_out
run() {
  echo "== $*"
  env "$@" python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['phase_ms_last_step']; print('ms/step %.2f conv %.2f lstm %.2f rvq %.2f' % (d['ms_per_step'], p['encoder_conv']+p['decoder_conv'], p['encoder_lstm']+p['decoder_lstm'], p['rvq']))"
}
run FCB_X=0
run FCB_TC_GROUP_MMAS=96
run FCB_TC_GROUP_MMAS=24
run FCB_TC_NA=4 FCB_TC_NB=2
run FCB_TC_NA=2 FCB_TC_NB=3
timeout 300 python -m pytest tests/test_gpu_layers.py -q -k tc 2>&1 | tail -2
