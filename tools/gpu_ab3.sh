#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
run() { echo "== $*"; env "$@" timeout 300 python bench.py --no-cpu --no-batch --steps 32 --warmup 4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('tok/s %.1f  ms/step %.3f  e2e %.1f  gemv %.0f GB/s (%.2f us/launch)' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['achieved'], r['avg_launch_us']))
    elif 'rror' in l: print(l.strip()[:300])
"; }
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_mulmat.py tests/test_gpu_llama_step.py -x -q 2>&1 | tail -2
for e in "GGML_B200_BS1_CLUSTER=1" "GGML_B200_BS1_CLUSTER=2" "GGML_B200_BS1_CLUSTER=4"; do echo "-- $e"; env $e timeout 300 python tools/bench_gemv.py --types q4_K,q6_K --cols 1 --shapes 4096x14336 --pdl 1 2>&1 | tail -2; done
run GGML_B200_BS1_CLUSTER=1
run GGML_B200_BS1_CLUSTER=2
run GGML_B200_BS1_CLUSTER=4
