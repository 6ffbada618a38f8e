#!/bin/bash
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" timeout 300 python bench.py --no-cpu --no-batch --steps 48 --warmup 4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('tok/s %.1f  ms/step %.3f  e2e %.1f  gemv %.0f GB/s (%.2f us/launch)' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['achieved'], r['avg_launch_us']))
    elif 'rror' in l: print(l.strip()[:300])
"; }
for i in 1 2; do
run GGML_B200_BS1_CLUSTER=4 GGML_B200_BS1_LPR32=1
run GGML_B200_BS1_CLUSTER=4 GGML_B200_BS1_LPR32=0
run GGML_B200_BS1_CLUSTER=1 GGML_B200_BS1_LPR32=0
done
