#!/bin/bash
# round-2 closing evidence on one B200: full GPU suite, both bench arms, smoke, stock ggml-cuda vs plugin through the same harness
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu_final.log; cat gpurun_out/r2_pytest_gpu_final.log
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 400 gpurun_out/r2_bench_reference.json
timeout 1500 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 1800 gpurun_out/r2_bench.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python tools/bench_configs.py > gpurun_out/r2_configs.jsonl 2> gpurun_out/r2_configs.err; cat gpurun_out/r2_configs.jsonl | cut -c1-400
bash tools/gpu_r2_refcuda.sh 2>&1 | grep "backend=" | sed 's/"ngl.*"prefill_ms"/"prefill_ms"/'
bash tools/gpu_r2_refcuda2.sh 2>&1 | grep "backend=" | sed 's/"ngl.*"prefill_ms"/"prefill_ms"/'
