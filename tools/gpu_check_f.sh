#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 300 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.log | cut -c1-300
for cfg in "--pdl 0 --l2pf 0" "--pdl 1 --l2pf 0" "--pdl 0 --l2pf 1" "--pdl 1 --l2pf 1"; do
  echo "== bench $cfg"; timeout 150 python bench.py --steps 32 --warmup 4 --no-cpu $cfg 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('tok/s %.1f  e2e %.1f  ms/step %.3f  launches/step %d  q4k gemv GB/s %.0f  whole-step GB/s %.0f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']/d['steps'], d['roofline']['achieved'], d['roofline']['whole_step']['achieved_gbs']))
except Exception as e: print('ERR', l[-600:])
"
done
echo "== ncu launches"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:b200_ -c 600 --csv --log-file $O/launches_r1f.csv python bench.py --steps 1 --warmup 3 --graphs 0 --no-cpu --pdl 0 > $O/ncu_bench.log 2>&1; echo "rc=$?"; python tools/summarize_launches.py $O/launches_r1f.csv
