#!/bin/bash
# round 2: small-batch mma kernel -- parity, micro-bench, batched bench leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mulmat.py -x -q -m gpu -k "small_batch or chunks_of_32 or batched_columns" 2>&1 | tail -15 > gpurun_out/r2l_tests.log
cat gpurun_out/r2l_tests.log
timeout 600 python tools/bench_gemv.py --types q4_K,q6_K,q4_0 --cols 8,16,32 --shapes 4096x4096,14336x4096,4096x14336,128256x4096 > gpurun_out/r2l_micro.log 2>&1
cat gpurun_out/r2l_micro.log
timeout 900 python bench.py --steps 32 --warmup 3 > gpurun_out/r2l_bench.log 2>&1
tail -3 gpurun_out/r2l_bench.log
