#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_fattn.py tests/test_gpu_llama_step.py tests/test_gpu_backend_ops.py -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 128 --warmup 8 --no-cpu > /tmp/b.json 2>/tmp/b.err; python - <<'PY'
import json
d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches/step',d['gpu_launches']/d['steps'],'frac',d['roofline']['frac'],'clocks',d['clocks'])
print('bs32',d['batched']['bs32_decode']['ms_per_step'],'pp',d['batched']['prefill_pp512']['ms_per_ubatch'])
PY
tail -2 /tmp/b.err
