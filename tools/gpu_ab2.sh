#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.log | cut -c1-400
run() { echo "== $*"; env "$@" timeout 300 python bench.py --no-cpu --steps 32 --warmup 4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('tok/s %.1f  ms/step %.3f  e2e %.1f  gemv %.0f GB/s (%.2f us/launch)' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['achieved'], r['avg_launch_us']))
    elif 'rror' in l: print(l.strip()[:300])
"; }
stamp "== micro"
for e in "GGML_B200_BS1_WARPS=32" "GGML_B200_BS1_WARPS=16"; do echo "-- $e"; env $e timeout 300 python tools/bench_gemv.py --types q4_K,q6_K --cols 1 --shapes 4096x4096,28672x4096,4096x14336,128256x4096 --pdl 1 2>&1 | tail -8; done
stamp "== step"
run GGML_B200_BS1_WARPS=32
run GGML_B200_BS1_WARPS=16
stamp done
