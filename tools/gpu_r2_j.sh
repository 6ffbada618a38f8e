#!/bin/bash
mkdir -p gpurun_out
for inf in 6 10 14 20 31; do
  GGML_B200_DS_INFLIGHT=$inf python bench.py --no-cpu --no-batch --steps 64 --warmup 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('inflight $inf: %.1f tok/s %.3f ms/step  gemv-only %.2f us/phase' % (d['value'], d['ms_per_step'], d['roofline']['avg_launch_us']))"
done
GGML_B200_DS_INFLIGHT=10 python tools/dstep_prof.py 4 0 2>&1 | tail -16
