#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mulmat.py -x -q -m gpu -k "mul_mat_id or small_batch" 2>&1 | tail -15
