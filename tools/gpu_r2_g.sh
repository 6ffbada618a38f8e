#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -4
python tools/batched_prof.py pp512 32 4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_mma_kernel -s 3 -c 1 -o gpurun_out/r2g_gemm_mma -f python tools/batched_prof.py pp512 2 1 > /dev/null 2>&1
