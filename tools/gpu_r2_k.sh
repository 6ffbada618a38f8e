#!/bin/bash
# per-launch path knobs
for env in "GGML_B200_BS1_CLUSTER=4" "GGML_B200_BS1_CLUSTER=2" "GGML_B200_BS1_CLUSTER=1" "GGML_B200_FA_VEC=1" "GGML_B200_FA_FUSED_COMBINE=1"; do
  env $env python bench.py --no-cpu --no-batch --steps 64 --warmup 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$env: %.1f tok/s %.3f ms/step  gemv-only %.2f us/launch frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['avg_launch_us'], d['roofline']['frac']))"
done
