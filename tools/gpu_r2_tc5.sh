#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_mulmat.py tests/test_gpu_gemm.py tests/test_gpu_llama_step.py -x -q 2>&1 | tail -4
timeout 400 python tools/bench_configs.py 2>&1 | tail -3
