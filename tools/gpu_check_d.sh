#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== repro prefill"; timeout 200 python tools/repro_prefill.py 512 2>&1 | tail -2
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 $O/pytest_gpu.log | cut -c1-300
for cfg in "--pdl 0 --l2pf 0 --fusion 1" "--pdl 0 --l2pf 0" "--pdl 0 --l2pf 1" "--pdl 1 --l2pf 0" "--pdl 1 --l2pf 1"; do
  echo "== bench $cfg"; timeout 300 python bench.py --steps 32 --warmup 4 --no-cpu $cfg 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('tok/s %.1f  e2e %.1f  ms/step %.3f  launches/step %d  q4k gemv GB/s %.0f  whole-step GB/s %.0f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']/d['steps'], d['roofline']['achieved'], d['roofline']['whole_step']['achieved_gbs']))
except Exception as e: print('ERR', l[-600:])
"
done
echo "== gemv micro"; timeout 300 python tools/bench_gemv.py --types q4_K,q6_K --cols 1 --shapes 4096x4096,14336x4096,4096x14336,128256x4096 2>&1 | tail -8
echo "== gemv micro 7 warps"; GGML_B200_GEMV_WARPS=7 timeout 300 python tools/bench_gemv.py --types q4_K,q6_K --cols 1 --shapes 4096x4096,14336x4096,4096x14336,128256x4096 2>&1 | tail -8
