#!/bin/bash
# closing evidence on one B200: GPU suite, both bench arms, smoke, configs #1/#2/#4 through the C ABI (outputs copied to profiles/r2_*)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2f_pytest_gpu.log; tail -2 gpurun_out/r2f_pytest_gpu.log
timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; tail -c 600 gpurun_out/r2f_bench_reference.json
timeout 1500 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 2500 gpurun_out/r2f_bench.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python tools/bench_configs.py > gpurun_out/r2f_configs.jsonl 2> gpurun_out/r2f_configs.err; cat gpurun_out/r2f_configs.jsonl | cut -c1-500
