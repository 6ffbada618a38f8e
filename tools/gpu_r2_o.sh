#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_mulmat.py tests/test_gpu_fattn.py tests/test_gpu_llama_step.py tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2o_tests.log
cat gpurun_out/r2o_tests.log
timeout 600 python tools/bench_gemv.py --types q4_K,q6_K,q4_0 --cols 16,32 --shapes 4096x4096,14336x4096,4096x14336,128256x4096 > gpurun_out/r2o_micro.log 2>&1
cat gpurun_out/r2o_micro.log
python tools/batched_prof.py bs32 32 8; python tools/batched_prof.py pp512 32 4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2o_bs32.csv python tools/batched_prof.py bs32 4 1 > gpurun_out/r2o_bs32.log 2>&1
python tools/summarize_launches.py gpurun_out/r2o_bs32.csv 2>&1 | grep -v "at::" | head -30
