#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_mulmat.py -x -q -m gpu -k "small_batch or mul_mat_id or chunks" 2>&1 | tail -4
python tools/bench_gemv.py --types q4_K,q6_K,q4_0 --cols 16,32 --shapes 4096x4096,14336x4096,4096x14336,128256x4096 --iters 10 2>&1 | cut -c1-150
python tools/batched_prof.py bs32 32 8
