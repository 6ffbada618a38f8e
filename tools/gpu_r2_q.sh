#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_mulmat.py tests/test_gpu_fattn.py tests/test_gpu_llama_step.py tests/test_gpu_glue.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2q_tests.log
cat gpurun_out/r2q_tests.log
python tools/batched_prof.py bs32 32 8
timeout 900 python bench.py --steps 32 --warmup 3 > gpurun_out/r2q_bench.log 2>&1
tail -1 gpurun_out/r2q_bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['batched'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2q_bs32.csv python tools/batched_prof.py bs32 4 1 > gpurun_out/r2q_bs32.log 2>&1
python tools/summarize_launches.py gpurun_out/r2q_bs32.csv 2>&1 | grep -v "at::" | head -30
