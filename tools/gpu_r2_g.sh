#!/bin/bash
echo "== gemv_mma (default)"; python tools/bench_gemv.py --types q4_K --cols 32 --shapes 4096x4096,14336x4096,4096x14336 --iters 10 2>&1 | cut -c1-150
echo "== gemm_mma TN=128 at M=32"; GGML_B200_MMA_MAX_M=4 python tools/bench_gemv.py --types q4_K --cols 32 --shapes 4096x4096,14336x4096,4096x14336 --iters 10 2>&1 | cut -c1-150
