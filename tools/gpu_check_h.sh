#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== gemm swap=0"; GGML_B200_GEMM_SWAP=0 timeout 200 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-250
echo "== gemm swap=1"; GGML_B200_GEMM_SWAP=1 timeout 200 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-250
