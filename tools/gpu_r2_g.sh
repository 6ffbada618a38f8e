#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_mulmat.py tests/test_gpu_llama_step.py -x -q -m gpu -k "small_batch or mul_mat_id or chunks or graph or replay" 2>&1 | tail -4
python tools/bench_gemv.py --types q4_K,q6_K --cols 32 --shapes 4096x4096,14336x4096,4096x14336,128256x4096 --iters 10 2>&1 | cut -c1-150
python tools/batched_prof.py bs32 32 8
export LD_LIBRARY_PATH=$PWD/cortex.llamacpp_b200:$PWD/oracle/_ref:$LD_LIBRARY_PATH
G=/tmp/rc_l3.gguf; python tools/make_gguf.py --model llama3-8b --ftype q4_k_m --out $G 2>/dev/null
GGML_BACKEND_PATH=$PWD/cortex.llamacpp_b200/libggml-b200.so LOGITS_DUMP_WARMUP=8 GGML_B200_GRAPHS=1 ./oracle/_ref/logits_dump $G - 99 512 128 f16 1 4 2>/dev/null | grep "^{"
