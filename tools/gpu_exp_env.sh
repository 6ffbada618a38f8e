#!/bin/bash
# env-knob sweep of the whole decode step (bench.py --no-cpu): one JSON line per configuration
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
run() { echo "== $*"; env "$@" timeout 300 python bench.py --no-cpu --steps 32 --warmup 4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('tok/s %.1f  ms/step %.3f  e2e %.1f  gemv %.0f GB/s (%.2f us/launch)' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['achieved'], r['avg_launch_us']))
    elif 'rror' in l: print(l.strip()[:300])
"; }
run A=0
run GGML_B200_GEMV_WARPS=9 GGML_B200_GEMV_SMEM_KB=110
run GGML_B200_GEMV_WARPS=9 GGML_B200_GEMV_SMEM_KB=100
run GGML_B200_GEMV_WARPS=6 GGML_B200_GEMV_SMEM_KB=74
run GGML_B200_GEMV_WARPS=12 GGML_B200_GEMV_SMEM_KB=110
run GGML_B200_PDL=0
