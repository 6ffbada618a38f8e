#!/bin/bash
# config #3 (32 slots, q8_0 KV) and #4 (Mixtral) : stock ggml-cuda vs the plugin through the same llama.cpp harness
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/cortex.llamacpp_b200:$PWD/oracle/_ref:/usr/local/cuda/lib64:$LD_LIBRARY_PATH
OUT=gpurun_out/r2_ref_cuda3.txt; : > $OUT
run() { # model ftype kv npar nprompt ngen
  G=/tmp/rc_$1_$2.gguf
  [ -f $G ] || python tools/make_gguf.py --model $1 --ftype $2 --out $G 2>/dev/null
  for be in cuda b200; do
    if [ $be = cuda ]; then export GGML_BACKEND_PATH=$PWD/oracle/_ref/cuda/libggml-cuda.so; else export GGML_BACKEND_PATH=$PWD/cortex.llamacpp_b200/libggml-b200.so; fi
    R=$(LOGITS_DUMP_WARMUP=4 GGML_B200_GRAPHS=1 timeout 900 ./oracle/_ref/logits_dump $G - 99 $5 $6 $3 $4 4 2>/tmp/err.txt | grep "^{" | tail -1)
    [ -z "$R" ] && R="FAILED: $(tail -3 /tmp/err.txt | tr '\n' ' ')"
    echo "$1 $2 kv=$3 n_parallel=$4 backend=$be $R" | tee -a $OUT
  done
}
run llama3-8b q4_k_m q8_0 32 64 32

run mixtral q4_k_m f16 1 512 32
