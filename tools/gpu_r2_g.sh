#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_mulmat.py tests/test_gpu_gemm.py -x -q -m gpu -k "mul_mat_id or gemm" 2>&1 | tail -5
timeout 900 python tools/bench_configs.py mixtral:q4_k_m 2>/dev/null | cut -c1-420
