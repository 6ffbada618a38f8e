#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_mma_kernel -s 4 -c 1 -o gpurun_out/r2p_mma -f python tools/bench_gemv.py --types q4_K --cols 32 --shapes 14336x4096 --iters 4 > gpurun_out/r2p_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:b200_fattn_kernel -s 2 -c 1 -o gpurun_out/r2p_fa -f python tools/batched_prof.py bs32 2 1 > gpurun_out/r2p_ncu2.log 2>&1
tail -2 gpurun_out/r2p_ncu.log gpurun_out/r2p_ncu2.log
