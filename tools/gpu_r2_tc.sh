#!/bin/bash
# gemm_tc bring-up: parity of the prompt-batch GEMM tests, then timing against gemm_mma / gemv_mma on the 8B shapes
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/r3a_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r3a_tests.log
SH=4096x4096,14336x4096,4096x14336
echo "== tc M=512"; timeout 200 python tools/bench_gemv.py --types q4_K,q6_K --cols 512 --shapes $SH --iters 10 2>&1 | tail -8
echo "== mma M=512"; GGML_B200_NO_GEMM_TC=1 timeout 200 python tools/bench_gemv.py --types q4_K,q6_K --cols 512 --shapes $SH --iters 10 2>&1 | tail -8
echo "== mma tn64 M=512"; GGML_B200_NO_GEMM_TC=1 GGML_B200_GEMM_TN=64 timeout 200 python tools/bench_gemv.py --types q4_K --cols 512 --shapes $SH --iters 10 2>&1 | tail -8
echo "== tc M=32"; GGML_B200_TC_MIN_M=5 timeout 200 python tools/bench_gemv.py --types q4_K,q6_K --cols 32 --shapes 1024x4096,$SH --iters 10 2>&1 | tail -10
echo "== gemv_mma M=32"; timeout 200 python tools/bench_gemv.py --types q4_K,q6_K --cols 32 --shapes 1024x4096,$SH --iters 10 2>&1 | tail -10
