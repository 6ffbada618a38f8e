#!/bin/bash
# gemm_tc integration check: full GPU suite, bench (no CPU legs), configs through the C ABI
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3d_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r3d_tests.log
timeout 600 python bench.py --no-cpu --steps 64 --warmup 4 > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r3d_bench.json') if x.startswith('{')]
j=json.loads(l[-1]); print('value', j['value'], 'frac', j['roofline']['frac']); print(json.dumps(j.get('batched'), indent=0)[:1500])
PY
timeout 600 python tools/bench_configs.py > gpurun_out/r3d_configs.jsonl 2>&1; tail -4 gpurun_out/r3d_configs.jsonl
