#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3f_tests.log
timeout 600 python bench.py --no-cpu --steps 64 --warmup 4 > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r3f_bench.json') if x.startswith('{')]
j=json.loads(l[-1]); print('value', j['value'], 'frac', j['roofline']['frac']); b=j['batched']; print('bs32', b['bs32_decode']['value'], b['bs32_decode']['ms_per_step'], 'pp512', b['prefill_pp512']['value'], b['prefill_pp512']['ms_per_ubatch'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r3f_pp512.csv python tools/prefill_prof.py 512 2 prefill > gpurun_out/r3f_pp512.log 2>&1
python tools/summarize_launches.py gpurun_out/r3f_pp512.csv 2>&1 | grep -i "fattn\|total"
