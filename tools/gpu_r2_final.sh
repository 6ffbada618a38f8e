#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu_final.log; cat gpurun_out/r2_pytest_gpu_final.log
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 600 gpurun_out/r2_bench_reference.json
timeout 1500 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 3500 gpurun_out/r2_bench.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
