#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -3
echo "== M=512 all"; timeout 100 python tools/bench_gemv.py --types q4_K,q6_K,q5_K --cols 512 --shapes 4096x4096,14336x4096,4096x14336 --iters 10 2>&1 | tail -9
echo "== M=32"; GGML_B200_TC_MIN_M=5 timeout 100 python tools/bench_gemv.py --types q4_K,q6_K --cols 32 --shapes 4096x4096,14336x4096,4096x14336 --iters 10 2>&1 | tail -6
echo "== real M=512"; GGML_B200_TC_PROF=1 timeout 100 python tools/bench_gemv.py --types q4_K --cols 512 --shapes 4096x4096 --iters 3 2>&1 | head -34
