#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -5
python tools/batched_prof.py pp512 32 4
