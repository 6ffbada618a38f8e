#!/bin/bash
mkdir -p gpurun_out
./tools/micro/imma_bench.bin 2>&1 | tee gpurun_out/r2n_imma.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_mma_kernel -s 4 -c 1 -o gpurun_out/r2n_mma -f python tools/bench_gemv.py --types q4_K --cols 32 --shapes 14336x4096 --iters 4 > gpurun_out/r2n_ncu.log 2>&1
tail -3 gpurun_out/r2n_ncu.log
