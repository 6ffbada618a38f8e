#!/bin/bash
# parity + A/B of the batch-1 GEMV kernel (gemv_bs1.cu) against the general kernel
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp "== pytest gpu (mulmat, step, glue)"; timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 $O/pytest_gpu.log | cut -c1-400
run() { echo "== $*"; env "$@" timeout 300 python bench.py --no-cpu --steps 32 --warmup 4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('tok/s %.1f  ms/step %.3f  e2e %.1f  gemv %.0f GB/s (%.2f us/launch)' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['achieved'], r['avg_launch_us']))
    elif 'rror' in l: print(l.strip()[:300])
"; }
stamp "== micro"
for e in "GGML_B200_BS1_OFF=1" "GGML_B200_BS1_CTAS=2" "GGML_B200_BS1_CTAS=1" "GGML_B200_BS1_CTAS=1 GGML_B200_BS1_SMEM_KB=220"; do echo "-- $e"; env $e timeout 300 python tools/bench_gemv.py --types q4_K,q6_K --cols 1 --shapes 4096x4096,28672x4096,4096x14336,128256x4096 --pdl ${PDL:-1} 2>&1 | tail -8; done
stamp "== step"
run GGML_B200_BS1_OFF=1
run GGML_B200_BS1_CTAS=2
run GGML_B200_BS1_CTAS=1
run GGML_B200_BS1_CTAS=1 GGML_B200_BS1_SMEM_KB=220
run GGML_B200_BS1_CTAS=2 GGML_B200_PDL=0
stamp done
